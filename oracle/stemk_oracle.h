/* oracle/stemk_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see stemk_oracle.c). */
#ifndef STEMK_ORACLE_H_
#define STEMK_ORACLE_H_
#include "../include/stemk.h"
#ifdef __cplusplus
extern "C" {
#endif
double oracle_pair(const stemk_params* p, const stemk_seqset_desc* X, uint32_t xi, const stemk_seqset_desc* Y,
                   uint32_t yi);
void oracle_pairs(const stemk_params* p, const stemk_seqset_desc* X, const stemk_seqset_desc* Y, size_t n_pairs,
                  const uint32_t* xi, const uint32_t* yi, double* out);
void oracle_gram(const stemk_params* p, const stemk_seqset_desc* S, int normalize, double* out);
void oracle_diag(const stemk_params* p, const stemk_seqset_desc* S, const uint32_t* sv_index, uint32_t n_sv,
                 double* out);
void oracle_cross(const stemk_params* p, const stemk_seqset_desc* T, const stemk_seqset_desc* S,
                  const uint32_t* sv_index, uint32_t n_sv, int normalize, double* out, double* self_out);
long oracle_print(const double* m, size_t rows, size_t cols, const int* labels, char* buf, long cap);
void oracle_pair_cost(const stemk_params* p, const stemk_seqset_desc* X, uint32_t xi, const stemk_seqset_desc* Y,
                      uint32_t yi, double* cells, double* flops);
/* BPLA / local-alignment kernel, bpla_kernel/bpla_kernel.cpp:16-175 */
void oracle_bpla_pairs(const stemk_bpla_params* p, const stemk_bpla_set* X, const stemk_bpla_set* Y, size_t n_pairs,
                       const uint32_t* xi, const uint32_t* yi, double* out);
/* BPLAKernel::compute_gradients, bpla_kernel/bpla_kernel.cpp:176-402: value[k] and grad[4k..4k+3] = d/d{alpha, beta, gap, ext} */
void oracle_bpla_gradients(const stemk_bpla_params* p, const stemk_bpla_set* X, const stemk_bpla_set* Y, size_t n_pairs,
                           const uint32_t* xi, const uint32_t* yi, double* value, double* grad);
/* naive stem kernel, stem_kernel/stem_kernel.cpp:282-351 (full_dp) with the base-pair classes of :353-420 */
void oracle_nstem_pairs(const stemk_nstem_params* p, const stemk_nstem_set* X, const stemk_nstem_set* Y, size_t n_pairs,
                        const uint32_t* xi, const uint32_t* yi, double* out);
/* band > 0: StemKernel::partial_dp with the band-only constraints (stem_kernel.cpp:14-83, 113-280); band == 0: full_dp */
void oracle_nstem_pairs_banded(const stemk_nstem_params* p, unsigned band, const stemk_nstem_set* X, const stemk_nstem_set* Y,
                               size_t n_pairs, const uint32_t* xi, const uint32_t* yi, double* out);
/* partial_dp under caller-supplied per-row windows: the c_low / c_high of alignment_constraints (stem_kernel.cpp:14-67) */
void oracle_nstem_pairs_windows(const stemk_nstem_params* p, const stemk_nstem_set* X, const stemk_nstem_set* Y, size_t n_pairs,
                                const uint32_t* xi, const uint32_t* yi, const uint32_t* win_off, const uint32_t* c_low,
                                const uint32_t* c_high, double* out);
#ifdef __cplusplus
}
#endif
#endif
