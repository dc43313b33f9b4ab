// Persistent warp-specialised fp16 GEMM on the 5th-generation tensor cores (sm_100a):
//   out[M,N] (+)= A[M,K] * W[N,K]^T (+ bias, activation / residual / row-remap epilogues)
// A and W are K-major fp16; accumulation is fp32 in TMEM.
//
//   warp 8      TMA producer   cp.async.bulk.tensor (128B swizzle) -> smem ring, mbarrier complete_tx
//   warp 9      MMA issuer     one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16),
//                              tcgen05.commit releases smem stages and publishes the TMEM accumulator
//   warps 0..7  epilogue       tcgen05.ld (32 lanes x 32 columns per warp-instruction) -> registers -> swizzled smem
//                              -> cp.async.bulk.tensor store (fp16) / cp.reduce.async.bulk.tensor add.f32 (residual)
//   warp 8      TMA producer,  warp 9  MMA issuer (see below)
// Two TMEM accumulator stages (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace bscnav {

enum GemmOut : int {
  kOutHalf = 0,        // fp16 out = acc + bias
  kOutHalfAct = 1,     // fp16 out = act(acc + bias)
  kOutResidual = 2,    // fp32 out += acc + bias           (residual stream)
  kOutPatchEmbed = 3,  // fp32 out[remap(m)] = acc + bias + table[m % np]   (patch embedding + position embedding)
  kOutFloat = 4,       // fp32 out = acc (+ bias)           (similarity matrices; qkv of the wide-range path)
  kOutSplitAct = 5,    // bf16 (hi | lo) out = split(act(acc + bias)): out[m, n] = hi, out[m, N + n] = lo   (wide-range path, fc1)
  kOutSims = 6,        // fp16 out = mask[n] ? acc : -inf, plus aux[m, n / 32] = max over each 32-column block (batched queries)
};

struct GemmEpilogue {
  void* out;
  const float* bias;    // [N] or null
  const float* table;   // kOutPatchEmbed: [np, N]
  long long ld_out;     // elements per output row
  int mode;
  int act;              // 0 = GELU(erf), 1 = QuickGELU
  int np, tokens, row_off;  // kOutPatchEmbed: out row = (m / np) * tokens + row_off + m % np
  int* range_flag;          // fp16 outputs: set to 1 when a value left the fp16 range (inf / nan after rounding); may be null
  const uint8_t* mask;      // kOutSims: [N] column participates (0 / 1)
  float* aux;               // kOutSims: [M, ld_aux] per-row maxima of 32-column blocks (fp32, before the fp16 rounding)
  long long ld_aux;
  // K slices (small M, gemm_f16_launch decides): scratch for the slices' fp32 partial tiles and one counter per output tile
  // (zero before the first launch; the kernel leaves them zero).  Null = never slice.
  float* kws;
  long long kws_floats;
  int* kcnt;
  int kcnt_len;
  int kslices;              // set by the launcher
};

// map_out: store map of the output tensor (make_store_map) for the fp16 / residual epilogues; null for the others
// map_w_half: the same weight matrix with a 128-row box (make_operand_map(..., 128)); when given (and N % 256 == 0, at least
// two row blocks) the 2-CTA cluster variant runs: each CTA loads half of the weight tile and multicasts it to its peer
// split != 0: wide-range operands ("bf16x3"): A is [M, 2K] and W is [N, 2K] bf16, each row holding the leading bf16 of every
// fp32 value followed by the bf16 of the remainder (hi | lo); the K loop runs three passes a_hi w_hi + a_lo w_hi + a_hi w_lo into the
// same fp32 accumulator (relative error ~2^-17 per product, fp32 range).  K is the logical depth, the maps cover 2K columns.
int gemm_f16_launch(const CUtensorMap* map_a, const CUtensorMap* map_w, const CUtensorMap* map_out, int M, int N, int K,
                    const GemmEpilogue& epi, cudaStream_t stream, const CUtensorMap* map_w_half = nullptr, int split = 0);
int make_store_map(CUtensorMap* map, const void* ptr, long long rows, long long N, bool is_float);

// 2-D K-major fp16 operand [rows, K] -> tensor map with box {64, box_rows}, 128-byte swizzle
int make_operand_map(CUtensorMap* map, const void* ptr, long long rows, long long K, int box_rows);

// tcgen05 attention (attention_tc.cu): map128 / map16 = make_operand_map over the qkv tensor with 128- and 16-row boxes
bool attention_tc_supported(int T);
int attention_tc_launch(const CUtensorMap* map128, const CUtensorMap* map16, __half* out, int batch, int T, int heads,
                        cudaStream_t stream);

// warp-specialised tcgen05 attention (attention_ws.cu): map128 / map16 = make_qkv_map3d with 128- and 16-row boxes
int make_qkv_map3d(CUtensorMap* map, const void* ptr, long long batch, long long T, long long width, int box_rows);
bool attention_ws_supported(int T);
int attention_ws_launch(const CUtensorMap* map128, const CUtensorMap* map16, __half* out, int batch, int T, int heads,
                        cudaStream_t stream);

// second-generation warp-specialised attention (attention_ws2.cu): P in TMEM, SIMT tail rows, TMA-store epilogue;
// map_out = make_out_map3d over the output tensor
int make_out_map3d(CUtensorMap* map, const void* ptr, long long batch, long long T, long long width);
bool attention_ws2_supported(int T);
int attention_ws2_launch(const CUtensorMap* map128, const CUtensorMap* map16, const CUtensorMap* map_out, const __half* qkv,
                         __half* out, int batch, int T, int heads, cudaStream_t stream, bool causal = false);

int gemm_block_n(int N);   // 256 when N % 256 == 0 (BN=128 is shared-memory-bandwidth bound at cta_group::1), else 128

}  // namespace bscnav
