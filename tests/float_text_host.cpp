// Host build of phylo2vec_b200/csrc/p2v_float_text.cuh for the CPU test suite (tests/test_float_text.py):
// the same source the CUDA kernels compile, called through ctypes.
#include "../phylo2vec_b200/csrc/p2v_float_text.cuh"

extern "C" {
int ft_format(double x, char* buf) { return p2v_ft::f64_to_shortest(x, buf); }
int ft_parse(const char* s, int len, double* out) { return p2v_ft::parse_f64(s, len, out); }
// batch forms so that millions of values are checked without ctypes call overhead
void ft_format_many(const double* x, long n, char* buf /* n * 32 */, int* lens) {
    for (long i = 0; i < n; ++i) lens[i] = p2v_ft::f64_to_shortest(x[i], buf + 32 * i);
}
// parse_f64, and parse_f64_exact for what it leaves undecided (rc2 counts how often that happened)
void ft_parse_full_many(const char* buf /* n * stride */, const int* lens, long n, int stride, double* out, int* rc, long* undecided) {
    uint32_t A[p2v_ft::FT_BIG_LIMBS], B[p2v_ft::FT_BIG_LIMBS];
    long und = 0;
    for (long i = 0; i < n; ++i) {
        rc[i] = p2v_ft::parse_f64(buf + (long)stride * i, lens[i], out + i);
        if (rc[i] == p2v_ft::PARSE_F64_UNDECIDED) { ++und; rc[i] = p2v_ft::parse_f64_exact(buf + (long)stride * i, lens[i], A, B, out + i); }
    }
    *undecided = und;
}
// parse_f64_exact on every literal (it must agree with parse_f64 wherever that one decides)
void ft_parse_exact_many(const char* buf, const int* lens, long n, int stride, double* out, int* rc) {
    uint32_t A[p2v_ft::FT_BIG_LIMBS], B[p2v_ft::FT_BIG_LIMBS];
    for (long i = 0; i < n; ++i) rc[i] = p2v_ft::parse_f64_exact(buf + (long)stride * i, lens[i], A, B, out + i);
}
void ft_parse_many(const char* buf /* n * stride */, const int* lens, long n, int stride, double* out, int* rc) {
    for (long i = 0; i < n; ++i) rc[i] = p2v_ft::parse_f64(buf + (long)stride * i, lens[i], out + i);
}
}
