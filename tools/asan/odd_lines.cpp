#include "pgio.h"
#include <cstdio>
#include <cstring>
#include <string>
static void wr(const char *fn, const std::string &s) { FILE *f = fopen(fn, "wb"); fwrite(s.data(), 1, s.size(), f); fclose(f); }
int main()
{
	pg_opt_t o; pgio_opt_init(&o);
	std::string good = "G1:P1\t100\t0\t100\t+\ts1#0#c1\t100000\t1000\t1500\t300\t300\t0\tAS:i:500\tms:i:500\tnp:i:100\tfs:i:0\tst:i:0\tcg:Z:50M200N50M\n";
	std::string cases[] = {
		"", "\n", "\n\n\n", "\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\t\n", "G7:P7\t10\n", "G7:P7", "0123456789abcdef\n", "0123456789abcde\tx\n", "0123456789abcdef\t\n",
		good, good + good, good.substr(0, good.size() - 1), std::string(good).replace(good.size() - 1, 1, "\r\n"),
		good.substr(0, 40) + std::string(1, '\0') + good.substr(41), good + "G2:P2\t50\t0\t50\t-\tctg\t1000\t10\t160\t150\t150\t0\tms:i:\tcg:Z:50M\n",
		good + "G3:P3\t50\t0\t50\t+\tctg\t1000\t10\t160\t150\t150\t0\tms:i:99999999999999999999\tcg:Z:50M\tzz:\tab\tms:\n",
		"G4:" + std::string(5000, 'x') + "\t50\t0\t50\t+\t" + std::string(3000, 'c') + "\t1000\t10\t160\t150\t150\t0\tms:i:77\tcg:Z:50M\n",
		good + "G5:P5\t60\t0\t60\t+\ts1#0#c1\t100000\t5000\t5177\t180\t180\t0\tms:i:280\tcg:Z:10M1F20M2G28M\n",
	};
	int k = 0;
	for (const std::string &c : cases) {
		char fn[64]; snprintf(fn, sizeof fn, "/tmp/asan/e%d.paf", k++);
		wr(fn, c);
		pg_data_t *d = pgio_data_init();
		int rc = pgio_read_paf(&o, d, fn);
		const char *two[2] = {fn, fn};
		pg_data_t *d2 = pgio_data_init();
		int n2 = pgio_read_pafs(&o, d2, two, 2, 2);
		printf("case %d: rc %d hits %ld; twice: %d files %ld hits, %d genes %d prots\n", k - 1, rc, (long)pgio_total_hits(d), n2, (long)pgio_total_hits(d2), d2->n_gene, d2->n_prot);
		pgio_write_bed(d2, 0, "/tmp/asan/e.bed");
		pgio_data_destroy(d); pgio_data_destroy(d2);
	}
	return 0;
}
