// Launchers of the backward-pass kernels (backward_kernels.cu); see dit_backward.cu for the sequence.
#pragma once
#include "kernels.h"

namespace b2 {

void bw_ln_bwd(const float* x, const float* du, const float* a, long long a_item_stride, int rows_per_item, float* dx,
               bool accumulate, float* rstd, float* mr, int M, int dim, float eps, cudaStream_t s);
// out[item][c] (+)= scale * sum over the item's rows of A[r, c] * B[r, c] * rs[r]   (B, rs optional)
void bw_colsum(const void* A, int a_dt, long long lda, const void* B, int b_dt, long long ldb, const float* rs, int items,
               int rows_per_item, int dim, float* out, long long out_item_stride, float scale, bool accumulate, float* scratch,
               cudaStream_t s);
size_t bw_colsum_scratch_bytes(int items, int rows_per_item, int dim);
void bw_axpy_gate(const float* xin, const float* y, const float* g, long long g_item_stride, int rows_per_item, float* xout,
                  long long M, int dim, cudaStream_t s);
void bw_mul_gate_cast(const float* dx, const float* g, long long g_item_stride, int rows_per_item, __half* out, long long ldo,
                      long long M, int dim, cudaStream_t s);
void bw_add(float* a, const float* b, long long n, cudaStream_t s);
void bw_fill(float* p, float v, long long n, cudaStream_t s);
void bw_scale_copy(const float* in, float* out, float sc, long long n, cudaStream_t s);
void bw_gelu_fwd(const __half* pre, __half* out, long long n, cudaStream_t s);
void bw_gelu_bwd(const __half* dh, const __half* pre, __half* out, long long n, cudaStream_t s);
void bw_rms_rope_fwd(const __half* raw, long long ld, const float* gamma, const float* cs, int rows_per_item, __half* out,
                     long long ldo, float* r_out, int M, int dim, float eps, cudaStream_t s);
void bw_rms_rope_bwd(const float* dout, const __half* raw, long long ld, const float* r, const float* gamma, const float* cs,
                     int rows_per_item, float* dun, __half* draw, long long ldd, int M, int dim, cudaStream_t s);
// softmax part of the attention backward for `heads` heads of one item: S, dP fp32 [heads][Lq128][lds] ->
// dS fp16 [heads][Lq128][ldk], dS^T and P^T fp16 [heads][Lk128][ldq]; stat: [heads * Lq128] float4 scratch
void bw_attn_softmax_bwd(const float* S, const float* dP, long long lds, int heads, int Lq, int Lq128, int Lk, int Lk128, int klen,
                         float scale, float* stat, __half* dS, long long ldk, __half* dST, __half* PT, long long ldq,
                         cudaStream_t s);
void bw_unpatchify_bwd(ItemPtrs dout, int B, int F, int Hp, int Wp, int out_dim, float scale, __half* dy16, float* dy32,
                       int rows_per_item, cudaStream_t s);
void bw_patchify_bwd(const float* dpatch, long long ld, int B, int C, int F, int H, int W, float scale, ItemPtrsMut dx,
                     int rows_per_item, cudaStream_t s);
void bw_small_fwd(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, bool silu_in, cudaStream_t s);
void bw_small_bwd(const float* dout, const float* W, const float* in, float* din, float* dW, float* db, int B, int K, int N,
                  bool silu_in, bool accumulate_din, float wscale, cudaStream_t s);   // dW, db += wscale * (...)
void bw_modtab_bwd(const float* dtab, int layers, int B, int dim, float* de0, float* dmod, float wscale, cudaStream_t s);
void bw_headtab_bwd(const float* dscale, const float* dshift, int B, int dim, float* dhead_mod, float* de, float wscale,
                    cudaStream_t s);

// ---- attn_bwd_tc.cu : fused FlashAttention backward (tcgen05 / TMEM), head_dim 128
struct AttnBwdParams {
  const __half* q; long long ldq;        // [items * Lq, ldq]   queries as the forward saw them, head h at columns h * 128
  const __half* k; long long ldk;        // [items * Lk, ldk]
  const __half* v; long long ldv;        // [items * Lk, ldv]   row-major values
  const float* O; long long ldo;         // fp32 [items * Lq, ldo]   the forward's output rows (AttnParams::out32)
  const __half* dO; long long lddo;      // [items * Lq, lddo]
  const float* lse;                      // [items][heads][Lq]  log2-domain row statistic of the forward (AttnParams::lse)
  float* dsum;                           // [items][heads][Lq]  scratch: rowsum(dO o O)
  float* dq; long long lddq;             // fp32 [items * Lq, lddq]   (zeroed and accumulated here)
  float* dk; long long lddk;             // fp32 [items * Lk, lddk]
  __half* dv; long long lddv;            // fp16 [items * Lk, lddv]
  int items, heads, Lq, Lk;
  int klen[MAX_ITEMS];
  float scale;
};
void launch_attention_backward(const AttnBwdParams& p, cudaStream_t stream);

}  // namespace b2
