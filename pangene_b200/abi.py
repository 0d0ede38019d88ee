"""ctypes mirrors of the pangene ABI structs (include/pangene_abi.h == reference pangene.h:19-124).

Only plain-old-data: the Python layer never interprets a hit itself, it hands pointers from the
reader (libpgio) to the hot-path library (libpgx) and on to the writers, exactly like the
reference's main() does (main.c:116-137).
"""
import ctypes as C


class pg_opt_t(C.Structure):
    _fields_ = [
        ("flag", C.c_uint32), ("gene_delim", C.c_int32),
        ("min_prot_ratio", C.c_double), ("min_prot_iden", C.c_double), ("score_adj_coef", C.c_double),
        ("min_ov_ratio", C.c_double), ("min_vertex_ratio", C.c_double),
        ("branch_diff", C.c_double), ("branch_diff_dist", C.c_double), ("branch_diff_cut", C.c_double),
        ("max_avg_occ", C.c_int32), ("max_degree", C.c_int32), ("max_dist_loci", C.c_int32),
        ("n_branch_flt", C.c_int32), ("min_arc_cnt", C.c_int32), ("local_dist", C.c_int32),
        ("local_count", C.c_int32),
        ("excl", C.c_void_p), ("incl", C.c_void_p), ("preferred", C.c_void_p),
    ]


class pg_exon_t(C.Structure):
    _fields_ = [("os", C.c_int32), ("oe", C.c_int32)]


class pg_prot_t(C.Structure):
    _fields_ = [("name", C.c_char_p), ("len", C.c_int32), ("gid", C.c_int32), ("rep", C.c_int32),
                ("n", C.c_int32), ("avg_score_adj", C.c_int32), ("max_score_ori", C.c_int32)]


class pg_gene_t(C.Structure):
    _fields_ = [("name", C.c_char_p), ("bits", C.c_uint32), ("rep_pid", C.c_int32)]


class pg_hit_t(C.Structure):
    _fields_ = [("pid", C.c_int32), ("qs", C.c_int32), ("qe", C.c_int32), ("cid", C.c_int32),
                ("mlen", C.c_int32), ("blen", C.c_int32), ("lof", C.c_int32), ("rank", C.c_int32),
                ("score_ori", C.c_int32), ("score_adj", C.c_int32), ("score_dom", C.c_int32),
                ("n_exon", C.c_int32), ("off_exon", C.c_int32), ("pid_dom", C.c_int32),
                ("pid_dom0", C.c_int32), ("bits", C.c_uint32),
                ("cs", C.c_int64), ("cm", C.c_int64), ("ce", C.c_int64)]


class pg_ctg_t(C.Structure):
    _fields_ = [("name", C.c_char_p), ("len", C.c_int64)]


class pg_genome_t(C.Structure):
    _fields_ = [("n_ctg", C.c_int32), ("m_ctg", C.c_int32), ("ctg", C.POINTER(pg_ctg_t)),
                ("n_hit", C.c_int32), ("m_hit", C.c_int32), ("hit", C.POINTER(pg_hit_t)),
                ("n_exon", C.c_int32), ("m_exon", C.c_int32), ("exon", C.POINTER(pg_exon_t)),
                ("label", C.c_char_p)]


class pg_data_t(C.Structure):
    _fields_ = [("d_ctg", C.c_void_p), ("d_gene", C.c_void_p), ("d_prot", C.c_void_p),
                ("n_genome", C.c_int32), ("m_genome", C.c_int32), ("genome", C.POINTER(pg_genome_t)),
                ("n_gene", C.c_int32), ("m_gene", C.c_int32), ("gene", C.POINTER(pg_gene_t)),
                ("n_prot", C.c_int32), ("m_prot", C.c_int32), ("prot", C.POINTER(pg_prot_t))]


class pg_seg_t(C.Structure):
    _fields_ = [("gid", C.c_int32), ("n_dom", C.c_int32), ("n_sub", C.c_int32), ("n_genome", C.c_int32),
                ("tot_cnt", C.c_int32), ("bits", C.c_uint32), ("n_dist_loci", C.c_int32 * 2)]


class pg_arc_t(C.Structure):
    _fields_ = [("x", C.c_uint64), ("n_genome", C.c_int32), ("tot_cnt", C.c_int32), ("avg_dist", C.c_int32),
                ("s1", C.c_int32), ("s2", C.c_int32), ("bits", C.c_uint32)]


class pg_graph_t(C.Structure):
    _fields_ = [("d", C.POINTER(pg_data_t)), ("g2s", C.POINTER(C.c_int32)),
                ("n_seg", C.c_int32), ("m_seg", C.c_int32), ("seg", C.POINTER(pg_seg_t)),
                ("n_arc", C.c_int32), ("m_arc", C.c_int32), ("arc", C.POINTER(pg_arc_t)),
                ("idx", C.POINTER(C.c_uint64))]


class pgx_packed_hit_t(C.Structure):
    """include/pgx_packed.h: the device-ready 48-byte record"""
    _fields_ = [(n, C.c_int32) for n in ("pid", "cid", "rank", "score_ori", "score_adj", "n_exon", "off_exon", "rev", "cs", "cm", "ce", "lof")]


class pgx_packed_t(C.Structure):
    _fields_ = [("n_genome", C.c_int32), ("m_genome", C.c_int32), ("n_hit", C.POINTER(C.c_int32)), ("m_hit", C.POINTER(C.c_int32)),
                ("hit", C.POINTER(C.POINTER(pgx_packed_hit_t)))]


class pgio_synth_t(C.Structure):
    _fields_ = [("shape", C.c_int32), ("n_genome", C.c_int32), ("genome_begin", C.c_int32),
                ("n_genome_total", C.c_int32), ("n_gene", C.c_int32), ("iso_per_gene", C.c_int32),
                ("n_ctg", C.c_int32), ("hits_per_genome", C.c_int32), ("seed", C.c_uint64),
                ("n_threads", C.c_int32)]


assert C.sizeof(pg_opt_t) == 128 and C.sizeof(pg_hit_t) == 88 and C.sizeof(pg_prot_t) == 32
assert C.sizeof(pg_gene_t) == 16 and C.sizeof(pg_seg_t) == 32 and C.sizeof(pg_arc_t) == 32
assert C.sizeof(pgx_packed_hit_t) == 48 and C.sizeof(pgx_packed_t) == 32
assert C.sizeof(pg_genome_t) == 56 and C.sizeof(pg_data_t) == 72 and C.sizeof(pg_graph_t) == 56

# option flag bits (pangene.h:8-17)
PG_F_FRAG_MODE = 0x20
PG_F_NO_JOINT_PSEUDO = 0x40
PG_F_ORI_FOR_BRANCH = 0x80
PG_F_CHECK_STRAND = 0x100
PG_F_DROP_SGL_EXON = 0x200

# bit-field word of pg_hit_t (include/pangene_abi.h PG_HB_*)
PG_HB_REV, PG_HB_FLT, PG_HB_FLT_SUB, PG_HB_FLT_OV, PG_HB_FLT_CHAIN = 0x001, 0x002, 0x004, 0x008, 0x010
PG_HB_PSEUDO, PG_HB_VTX, PG_HB_SHADOW, PG_HB_REP, PG_HB_WBR_SHIFT, PG_HB_WBR_MASK = 0x020, 0x040, 0x080, 0x100, 9, 0x600
