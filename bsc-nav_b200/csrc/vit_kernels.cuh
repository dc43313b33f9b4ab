// Host-callable launchers of the ViT support kernels (vit_kernels.cu).
#pragma once
#include "common.cuh"

namespace bscnav {

// tables: preprocess_table_bytes(S) bytes of device scratch for the resize tap tables (null: per-pixel kernel)
int64_t preprocess_table_bytes(int S);
int preprocess_frames(const uint8_t* rgb, int batch, int H, int W, int ch, int S, int P, int k_pad, __half* a0, float* out_pre,
                      float* tables, cudaStream_t stream);
int im2col_images(const float* img, int batch, int S, int P, int k_pad, __half* a0, cudaStream_t stream);
int fill_prefix_tokens(float* x, const float* cls_pos, const float* reg, int batch, int tokens, int n_reg, int D, cudaStream_t stream);
int layernorm_half(const float* x, const float* g, const float* b, float eps, long long rows, int D, __half* out, cudaStream_t s);
int layernorm_inplace(float* x, const float* g, const float* b, float eps, long long rows, int D, cudaStream_t s);
int layernorm_patch_tokens(const float* x, const float* g, const float* b, float eps, int batch, int tokens, int skip, int D,
                           float* out, cudaStream_t s);
// mma.sync kernel, head dim 64 | 80; causal (keys <= query row, text tower): head dim 64
int attention_f16(const __half* qkv, __half* out, int batch, int T, int heads, int head_dim, bool causal, cudaStream_t stream);
// image-embedding head (CLIP encode_image): LayerNorm of each frame's cls row, times proj [D, E] (fp32) -> out [batch, E]
// ids != null: pool the row at argmax(ids[b, :]) (end-of-text token of the text tower) instead of row 0
int cls_embed(const float* x, const float* g, const float* b, float eps, int batch, int tokens, int D, const float* proj, int E,
              float* out, const int32_t* ids, cudaStream_t stream);
// text tower input: x[b, t, :] = tok_emb[ids[b, t]] + pos[t]
int text_embed(const int32_t* ids, const float* tok_emb, int vocab, const float* pos, int batch, int ctx, int D, float* x,
               cudaStream_t stream);

// wide-range ("bf16x3") path, wide_kernels.cu: operands as (hi | lo) bf16 row pairs, fp32 attention (head dim 64)
int split_rows(const float* src, long long rows, int K, void* dst, cudaStream_t stream);
int im2col_images_split(const float* img, int batch, int S, int P, int k_pad, void* a0, cudaStream_t stream);
int layernorm_split(const float* x, const float* g, const float* b, float eps, long long rows, int D, void* out, cudaStream_t stream);
int attention_f32_split(const float* qkv, void* out, int batch, int T, int heads, bool causal, cudaStream_t stream);

}  // namespace bscnav
