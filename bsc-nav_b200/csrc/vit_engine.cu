// ViT engine: owns nothing but TMA descriptors; weights and workspace are caller-owned device buffers.
// Replaces dinov2.forward_features (memory_2.py:591-592, 732-742).  Weight blob layout (see vit.py:pack_weights):
//   half blob : patch_w [D, k_pad] | per layer { qkv_w [3D, D] (q rows pre-scaled by 1/sqrt(dh)), proj_w [D, D] (LayerScale folded),
//               fc1_w [M, D], fc2_w [D, M] (LayerScale folded) }
//   float blob: patch_b [D] | cls_pos [D] | reg [n_reg, D] | pos [Np, D] | ln_pre_w, ln_pre_b [D] (if ln_pre) |
//               per layer { ln1_w, ln1_b [D], qkv_b [3D], proj_b [D], ln2_w, ln2_b [D], fc1_b [M], fc2_b [D] } | ln_post_w, ln_post_b [D]
// Text tower (cfg.text_context > 0, open_clip encode_text for the agent's goal / label matching, BSCAgent.py:541-557, 708-711):
//   half blob : per layer { qkv_w, proj_w, fc1_w, fc2_w };  float blob: pos [ctx, D] | per layer { ... } | ln_final_w, ln_final_b
//   causal attention, no patch embedding / cls / registers / pre-LN; the token-embedding table is passed per call.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include "gemm_tcgen05.cuh"
#include "vit_kernels.cuh"

namespace bscnav {

struct LayerOff {
  long long qkv_w, proj_w, fc1_w, fc2_w;                                   // half blob (elements)
  long long ln1_w, ln1_b, qkv_b, proj_b, ln2_w, ln2_b, fc1_b, fc2_b;       // float blob (elements)
};

struct VitLayout {
  long long patch_w = 0, half_total = 0;
  long long patch_b = 0, cls_pos = 0, reg = 0, pos = 0, ln_pre_w = 0, ln_pre_b = 0, ln_post_w = 0, ln_post_b = 0, float_total = 0;
  LayerOff layer[64];
};

static int vit_layout(const bscnav_vit_cfg_t* c, VitLayout* L) {
  if (!c || c->depth <= 0 || c->depth > 64 || c->dim % 128 || c->mlp % 128 || c->heads <= 0) return BSCNAV_EINVAL;
  const int hd = c->head_dim ? c->head_dim : 64;
  if ((hd != 64 && hd != 80) || c->dim != c->heads * hd) return BSCNAV_EINVAL;
  const bool text = c->text_context > 0;
  if (c->text_context < 0 || (text && (hd != 64 || c->n_reg || c->ln_pre))) return BSCNAV_EINVAL;
  if (!text && (c->patch <= 0 || c->image % c->patch || c->patch_k_pad % 64 || c->patch_k_pad < 3 * c->patch * c->patch)) return BSCNAV_EINVAL;
  if (c->precision != 0 && c->precision != 1) return BSCNAV_EINVAL;
  if (c->precision == 1 && hd != 64) return BSCNAV_EINVAL;              // the fp32 attention kernel: head dim 64
  const long long D = c->dim, M = c->mlp;
  const long long ws = c->precision == 1 ? 2 : 1;                       // wide-range path: (hi | lo) pairs, rows of 2K
  const long long Np = text ? c->text_context : (long long)(c->image / c->patch) * (c->image / c->patch);
  long long h = 0, f = 0;
  if (!text) {
    L->patch_w = h; h += ws * D * c->patch_k_pad;
    L->patch_b = f; f += D;
    L->cls_pos = f; f += D;
    L->reg = f; f += (long long)c->n_reg * D;
  }
  L->pos = f; f += Np * D;
  if (c->ln_pre) { L->ln_pre_w = f; f += D; L->ln_pre_b = f; f += D; }
  for (int i = 0; i < c->depth; ++i) {
    LayerOff& o = L->layer[i];
    o.qkv_w = h; h += ws * 3 * D * D;
    o.proj_w = h; h += ws * D * D;
    o.fc1_w = h; h += ws * M * D;
    o.fc2_w = h; h += ws * D * M;
    o.ln1_w = f; f += D; o.ln1_b = f; f += D;
    o.qkv_b = f; f += 3 * D;
    o.proj_b = f; f += D;
    o.ln2_w = f; f += D; o.ln2_b = f; f += D;
    o.fc1_b = f; f += M;
    o.fc2_b = f; f += D;
  }
  L->ln_post_w = f; f += D;
  L->ln_post_b = f; f += D;
  L->half_total = h;
  L->float_total = f;
  return BSCNAV_OK;
}

struct Workspace {
  __half* a0;    // [B*Np, k_pad]  patch rows
  float* x;      // [B*T, D]       residual stream
  __half* y;     // [B*T, D]       LayerNorm output / attention output
  __half* qkv;   // [B*T, 3D]
  __half* ao;    // [B*T, D]
  __half* hmid;  // [B*T, M]
  // wide-range path: a0 / y / ao / hmid hold (hi | lo) bf16 rows of twice the width, qkv is fp32
  __half* a0h;   // fp16 patch rows written by the pre-processing kernel (scratch; the GEMM reads the split rows in a0)
  float* pre_img; // [B, 3, S, S] pre-processed frames (input of the split im2col)
  float* pre_tab; // resize tap tables of the pre-processing kernel
  int* range_flag; // set by the GEMM epilogues when an fp16 activation overflowed
  float* kws;      // K-sliced GEMMs of small batches: partial tiles (gemm_tcgen05.cuh), kKwsFloats floats
  int* kcnt;       // and their tile counters (kKcntLen, zero between launches)
  long long total;
};
static constexpr long long kKwsFloats = 152LL * 128 * 128;   // one 128 x 128 fp32 partial tile per SM and a little slack
static constexpr int kKcntLen = 256;
static void workspace_layout(const bscnav_vit_cfg_t* c, int max_batch, Workspace* w, char* base) {
  const bool text = c->text_context > 0;
  const long long D = c->dim, M = c->mlp, G = text ? 0 : c->image / c->patch, Np = G * G, B = max_batch;
  const long long T = text ? c->text_context : 1 + c->n_reg + Np;
  long long off = 0;
  auto take = [&](long long bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 1024); return p; };
  const long long wide = c->precision == 1 ? 2 : 1;
  w->a0 = reinterpret_cast<__half*>(take(text ? 0 : wide * B * Np * c->patch_k_pad * 2));
  w->x = reinterpret_cast<float*>(take(B * T * D * 4));
  w->y = reinterpret_cast<__half*>(take(wide * B * T * D * 2));
  w->qkv = reinterpret_cast<__half*>(take(wide * B * T * 3 * D * 2));      // wide: fp32
  w->ao = reinterpret_cast<__half*>(take(wide * B * T * D * 2));
  w->hmid = reinterpret_cast<__half*>(take(wide * B * T * M * 2));
  w->a0h = reinterpret_cast<__half*>(take(text || wide == 1 ? 0 : B * Np * c->patch_k_pad * 2));
  w->pre_img = reinterpret_cast<float*>(take(text || wide == 1 ? 0 : B * 3 * (long long)c->image * c->image * 4));
  w->pre_tab = reinterpret_cast<float*>(take(text ? 0 : preprocess_table_bytes(c->image)));
  w->range_flag = reinterpret_cast<int*>(take(256));
  w->kws = reinterpret_cast<float*>(take(kKwsFloats * 4));
  w->kcnt = reinterpret_cast<int*>(take(kKcntLen * 4));
  w->total = off;
}

struct VitEngine {
  bscnav_vit_cfg_t cfg;
  VitLayout L;
  Workspace ws;
  const __half* wh;
  const float* wf;
  int max_batch, Np, T;
  CUtensorMap map_a0, map_y, map_ao, map_h;          // activation operands
  CUtensorMap map_patch_w, half_patch_w;             // half_: 128-row boxes (CTA-pair GEMM, small-M 128-wide tiles)
  bool has_half_patch = false;
  CUtensorMap out_qkv, out_h, out_x;                 // store maps of the GEMM outputs
  CUtensorMap qkv128, qkv16;                         // qkv tensor as attention operand (128- and 16-row boxes)
  CUtensorMap qkv3d128, qkv3d16;                     // the same tensor as {column, token, frame} (attention_ws.cu)
  CUtensorMap ao3d;                                  // attention output as a 3-D store map (attention_ws2.cu)
  int attn_impl = 0;                                 // 0 mma.sync, 1 tcgen05 v1, 2 tcgen05 warp-specialised, 3 second generation
  CUtensorMap map_qkv[64], map_proj[64], map_fc1[64], map_fc2[64];
  CUtensorMap half_qkv[64], half_proj[64], half_fc1[64], half_fc2[64];   // 128-row boxes of the same weights (cluster multicast)
  bool has_half[4] = {false, false, false, false};
  bool a0_cleared = false;
  // optional per-kernel timing (bench.py's roofline leg): CUDA events around every launch on the caller's stream
  bool profile = false;
  struct Span { int cat; double work; cudaEvent_t e0, e1; };
  std::vector<Span> spans;
  // CUDA graphs of the transformer blocks for small batches (one frame per call: the 168 launches of a forward cost the host
  // 1.4 ms to enqueue, more than the kernels take): batch -> {executable graph or null (= run eagerly), launches in it}
  struct BlockGraph { cudaGraphExec_t exec; int launches; bool warmed; };
  std::map<int, BlockGraph> graphs;
  cudaStream_t capture_stream = nullptr;
  ~VitEngine() {
    for (auto& g : graphs) if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
    if (capture_stream) cudaStreamDestroy(capture_stream);
  }
};

enum ProfCat : int { kCatGemm = 0, kCatAttention = 1, kCatLayerNorm = 2, kCatPre = 3, kCatCount = 4 };

struct ProfScope {
  VitEngine* e; cudaStream_t s; size_t idx;
  ProfScope(VitEngine* e_, cudaStream_t s_, int cat, double work) : e(e_), s(s_), idx(0) {
    if (!e->profile) return;
    VitEngine::Span sp{cat, work, nullptr, nullptr};
    cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, s);
    e->spans.push_back(sp);
    idx = e->spans.size();
  }
  ~ProfScope() { if (idx) cudaEventRecord(e->spans[idx - 1].e1, s); }
};

static int run_layers(VitEngine* e, int batch, cudaStream_t stream);
static int run_blocks(VitEngine* e, int batch, cudaStream_t stream);

// out_tokens: LayerNorm(post) patch tokens [batch, Np, D]; or (proj != null) out_embed: the cls token's image embedding [batch, E]
static int vit_run(VitEngine* e, int batch, float* out_tokens, cudaStream_t stream, const float* proj = nullptr, int embed_dim = 0,
                   float* out_embed = nullptr) {
  const bscnav_vit_cfg_t& c = e->cfg;
  if (c.text_context > 0) return BSCNAV_EINVAL;        // a text engine has no image path
  const int D = c.dim, T = e->T, Np = e->Np;
  const int rows = batch * T;
  const float* wf = e->wf;
  // patch embedding + position embedding, written straight into the residual stream at the patch rows
  GemmEpilogue ep{};
  ep.out = e->ws.x; ep.bias = wf + e->L.patch_b; ep.table = wf + e->L.pos; ep.ld_out = D; ep.mode = kOutPatchEmbed;
  ep.np = Np; ep.tokens = T; ep.row_off = 1 + c.n_reg;
  { ProfScope ps(e, stream, kCatGemm, 2.0 * batch * Np * (double)D * c.patch_k_pad);
    if (int r = gemm_f16_launch(&e->map_a0, &e->map_patch_w, nullptr, batch * Np, D, c.patch_k_pad, ep, stream,
                                e->has_half_patch ? &e->half_patch_w : nullptr, c.precision)) return r; }
  if (int r = fill_prefix_tokens(e->ws.x, wf + e->L.cls_pos, wf + e->L.reg, batch, T, c.n_reg, D, stream)) return r;
  if (c.ln_pre)
    if (int r = layernorm_inplace(e->ws.x, wf + e->L.ln_pre_w, wf + e->L.ln_pre_b, c.ln_eps, rows, D, stream)) return r;
  if (int r = run_blocks(e, batch, stream)) return r;
  if (proj) return cls_embed(e->ws.x, wf + e->L.ln_post_w, wf + e->L.ln_post_b, c.ln_eps, batch, T, D, proj, embed_dim, out_embed, nullptr, stream);
  return layernorm_patch_tokens(e->ws.x, wf + e->L.ln_post_w, wf + e->L.ln_post_b, c.ln_eps, batch, T, 1 + c.n_reg, D,
                                out_tokens, stream);
}

// the transformer blocks over the residual stream ws.x [batch*T, D] (image and text towers)
static int run_layers_wide(VitEngine* e, int batch, cudaStream_t stream);

static int run_layers(VitEngine* e, int batch, cudaStream_t stream) {
  const bscnav_vit_cfg_t& c = e->cfg;
  if (c.precision == 1) return run_layers_wide(e, batch, stream);
  const int D = c.dim, M = c.mlp, T = e->T;
  const int rows = batch * T;
  const float* wf = e->wf;
  const bool causal = c.text_context > 0;
  for (int i = 0; i < c.depth; ++i) {
    const LayerOff& o = e->L.layer[i];
    { ProfScope ps(e, stream, kCatLayerNorm, (double)rows * D * 6);
      if (int r = layernorm_half(e->ws.x, wf + o.ln1_w, wf + o.ln1_b, c.ln_eps, rows, D, e->ws.y, stream)) return r; }
    GemmEpilogue e1{};
    e1.kws = e->ws.kws; e1.kws_floats = kKwsFloats; e1.kcnt = e->ws.kcnt; e1.kcnt_len = kKcntLen;
    e1.out = e->ws.qkv; e1.bias = wf + o.qkv_b; e1.ld_out = 3 * D; e1.mode = kOutHalf; e1.range_flag = e->ws.range_flag;
    { ProfScope ps(e, stream, kCatGemm, 2.0 * rows * 3.0 * D * D);
      if (int r = gemm_f16_launch(&e->map_y, &e->map_qkv[i], &e->out_qkv, rows, 3 * D, D, e1, stream, e->has_half[0] ? &e->half_qkv[i] : nullptr)) return r; }
    { ProfScope ps(e, stream, kCatAttention, 4.0 * batch * (double)T * T * D);
      if (e->attn_impl == 3 && batch == 1 && !causal) {
        // one frame: 16 (frame, head) pairs leave the persistent tcgen05 kernel one CTA each on 16 SMs; the mma.sync kernel spreads
        // the row-group rounds of every pair over three CTAs (1.432 -> 1.394 ms per ViT-L/14 forward; slower from B = 4 on)
        if (int r = attention_f16(e->ws.qkv, e->ws.ao, batch, T, c.heads, 64, false, stream)) return r;
      } else if (e->attn_impl == 3) {
        if (int r = attention_ws2_launch(&e->qkv3d128, &e->qkv3d16, &e->ao3d, e->ws.qkv, e->ws.ao, batch, T, c.heads, stream, causal)) return r;
      } else if (e->attn_impl == 2) {
        // the 3-D map's frame extent is the engine's max_batch: frames >= batch are simply never addressed
        if (int r = attention_ws_launch(&e->qkv3d128, &e->qkv3d16, e->ws.ao, batch, T, c.heads, stream)) return r;
      } else if (e->attn_impl == 1) { if (int r = attention_tc_launch(&e->qkv128, &e->qkv16, e->ws.ao, batch, T, c.heads, stream)) return r; }
      else if (int r = attention_f16(e->ws.qkv, e->ws.ao, batch, T, c.heads, c.head_dim ? c.head_dim : 64, causal, stream)) return r; }
    GemmEpilogue e2{};
    e2.kws = e->ws.kws; e2.kws_floats = kKwsFloats; e2.kcnt = e->ws.kcnt; e2.kcnt_len = kKcntLen;
    e2.out = e->ws.x; e2.bias = wf + o.proj_b; e2.ld_out = D; e2.mode = kOutResidual;
    { ProfScope ps(e, stream, kCatGemm, 2.0 * rows * (double)D * D);
      if (int r = gemm_f16_launch(&e->map_ao, &e->map_proj[i], &e->out_x, rows, D, D, e2, stream, e->has_half[1] ? &e->half_proj[i] : nullptr)) return r; }
    { ProfScope ps(e, stream, kCatLayerNorm, (double)rows * D * 6);
      if (int r = layernorm_half(e->ws.x, wf + o.ln2_w, wf + o.ln2_b, c.ln_eps, rows, D, e->ws.y, stream)) return r; }
    GemmEpilogue e3{};
    e3.kws = e->ws.kws; e3.kws_floats = kKwsFloats; e3.kcnt = e->ws.kcnt; e3.kcnt_len = kKcntLen;
    e3.out = e->ws.hmid; e3.bias = wf + o.fc1_b; e3.ld_out = M; e3.mode = kOutHalfAct; e3.act = c.act; e3.range_flag = e->ws.range_flag;
    { ProfScope ps(e, stream, kCatGemm, 2.0 * rows * (double)M * D);
      if (int r = gemm_f16_launch(&e->map_y, &e->map_fc1[i], &e->out_h, rows, M, D, e3, stream, e->has_half[2] ? &e->half_fc1[i] : nullptr)) return r; }
    GemmEpilogue e4{};
    e4.kws = e->ws.kws; e4.kws_floats = kKwsFloats; e4.kcnt = e->ws.kcnt; e4.kcnt_len = kKcntLen;
    e4.out = e->ws.x; e4.bias = wf + o.fc2_b; e4.ld_out = D; e4.mode = kOutResidual;
    { ProfScope ps(e, stream, kCatGemm, 2.0 * rows * (double)M * D);
      if (int r = gemm_f16_launch(&e->map_h, &e->map_fc2[i], &e->out_x, rows, D, M, e4, stream, e->has_half[3] ? &e->half_fc2[i] : nullptr)) return r; }
  }
  return BSCNAV_OK;
}

// run_layers, replayed from a CUDA graph when the batch is small.  The blocks read and write only the engine's workspace and
// weights, so a graph captured once per batch size stays valid; the first call of a batch size runs eagerly (function attributes,
// module loading), the second is captured -- on a stream of the engine's own, since the caller's may be the legacy default stream,
// which cannot be captured -- and launched into the caller's stream.  BSCNAV_GRAPH=0 disables it; so does the profiling hook; a
// caller that is itself capturing gets the plain launches recorded into its own graph.
static constexpr int kGraphMaxRows = 4096;             // larger batches are bound by the kernels, not by the host's launch rate
static int run_blocks(VitEngine* e, int batch, cudaStream_t stream) {
  static int use_graph = -1;
  if (use_graph < 0) { const char* v = getenv("BSCNAV_GRAPH"); use_graph = v ? atoi(v) : 1; }
  if (!use_graph || e->profile || batch * e->T > kGraphMaxRows) return run_layers(e, batch, stream);
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) { cudaGetLastError(); return run_layers(e, batch, stream); }
  auto it = e->graphs.find(batch);
  if (it == e->graphs.end()) {
    e->graphs[batch] = VitEngine::BlockGraph{nullptr, 0, true};
    return run_layers(e, batch, stream);
  }
  VitEngine::BlockGraph& bg = it->second;
  if (bg.warmed && !bg.exec) {
    bg.warmed = false;                                  // one attempt: on any failure this batch size stays eager
    if (!e->capture_stream && cudaStreamCreateWithFlags(&e->capture_stream, cudaStreamNonBlocking) != cudaSuccess) {
      cudaGetLastError(); e->capture_stream = nullptr; return run_layers(e, batch, stream);
    }
    const long long l0 = g_launches.load(std::memory_order_relaxed);
    const cudaError_t cb = cudaStreamBeginCapture(e->capture_stream, cudaStreamCaptureModeThreadLocal);
    if (cb != cudaSuccess) {
      fprintf(stderr, "bscnav: graph capture not started (%s); batch %d stays eager\n", cudaGetErrorString(cb), batch);
      cudaGetLastError();
      return run_layers(e, batch, stream);
    }
    const int r = run_layers(e, batch, e->capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->capture_stream, &graph);
    const int n = (int)(g_launches.load(std::memory_order_relaxed) - l0);
    g_launches.fetch_sub(n, std::memory_order_relaxed);  // enqueued into the graph, not run: each replay counts them
    cudaGraphExec_t exec = nullptr;
    cudaError_t ci = cudaSuccess;
    if (r == BSCNAV_OK && ce == cudaSuccess && graph && (ci = cudaGraphInstantiate(&exec, graph, 0)) == cudaSuccess) { bg.exec = exec; bg.launches = n; }
    else {
      fprintf(stderr, "bscnav: graph of the blocks at batch %d not built (run %d, end capture: %s, instantiate: %s); stays eager\n", batch, r,
              cudaGetErrorString(ce), cudaGetErrorString(ci));
      cudaGetLastError();
    }
    if (graph) cudaGraphDestroy(graph);
  }
  if (!bg.exec) return run_layers(e, batch, stream);
  BSC_CUDA(cudaGraphLaunch(bg.exec, stream));
  g_launches.fetch_add(bg.launches, std::memory_order_relaxed);
  return BSCNAV_OK;
}

// the same blocks on the wide-range path: split bf16 operands, three GEMM passes, fp32 qkv and attention
static int run_layers_wide(VitEngine* e, int batch, cudaStream_t stream) {
  const bscnav_vit_cfg_t& c = e->cfg;
  const int D = c.dim, M = c.mlp, T = e->T;
  const int rows = batch * T;
  const float* wf = e->wf;
  const bool causal = c.text_context > 0;
  for (int i = 0; i < c.depth; ++i) {
    const LayerOff& o = e->L.layer[i];
    { ProfScope ps(e, stream, kCatLayerNorm, (double)rows * D * 8);
      if (int r = layernorm_split(e->ws.x, wf + o.ln1_w, wf + o.ln1_b, c.ln_eps, rows, D, e->ws.y, stream)) return r; }
    GemmEpilogue e1{};
    e1.out = e->ws.qkv; e1.bias = wf + o.qkv_b; e1.ld_out = 3 * D; e1.mode = kOutFloat;
    { ProfScope ps(e, stream, kCatGemm, 3 * 2.0 * rows * 3.0 * D * D);
      if (int r = gemm_f16_launch(&e->map_y, &e->map_qkv[i], nullptr, rows, 3 * D, D, e1, stream, e->has_half[0] ? &e->half_qkv[i] : nullptr, 1)) return r; }
    { ProfScope ps(e, stream, kCatAttention, 4.0 * batch * (double)T * T * D);
      if (int r = attention_f32_split(reinterpret_cast<const float*>(e->ws.qkv), e->ws.ao, batch, T, c.heads, causal, stream)) return r; }
    GemmEpilogue e2{};
    e2.out = e->ws.x; e2.bias = wf + o.proj_b; e2.ld_out = D; e2.mode = kOutResidual;
    { ProfScope ps(e, stream, kCatGemm, 3 * 2.0 * rows * (double)D * D);
      if (int r = gemm_f16_launch(&e->map_ao, &e->map_proj[i], &e->out_x, rows, D, D, e2, stream, e->has_half[1] ? &e->half_proj[i] : nullptr, 1)) return r; }
    { ProfScope ps(e, stream, kCatLayerNorm, (double)rows * D * 8);
      if (int r = layernorm_split(e->ws.x, wf + o.ln2_w, wf + o.ln2_b, c.ln_eps, rows, D, e->ws.y, stream)) return r; }
    GemmEpilogue e3{};
    e3.out = e->ws.hmid; e3.bias = wf + o.fc1_b; e3.ld_out = 2 * M; e3.mode = kOutSplitAct; e3.act = c.act;
    { ProfScope ps(e, stream, kCatGemm, 3 * 2.0 * rows * (double)M * D);
      if (int r = gemm_f16_launch(&e->map_y, &e->map_fc1[i], nullptr, rows, M, D, e3, stream, e->has_half[2] ? &e->half_fc1[i] : nullptr, 1)) return r; }
    GemmEpilogue e4{};
    e4.out = e->ws.x; e4.bias = wf + o.fc2_b; e4.ld_out = D; e4.mode = kOutResidual;
    { ProfScope ps(e, stream, kCatGemm, 3 * 2.0 * rows * (double)M * D);
      if (int r = gemm_f16_launch(&e->map_h, &e->map_fc2[i], &e->out_x, rows, D, M, e4, stream, e->has_half[3] ? &e->half_fc2[i] : nullptr, 1)) return r; }
  }
  return BSCNAV_OK;
}

}  // namespace bscnav

using namespace bscnav;

extern "C" {

int64_t bscnav_vit_half_elems(const bscnav_vit_cfg_t* c) {
  VitLayout L;
  if (vit_layout(c, &L)) return BSCNAV_EINVAL;
  return L.half_total;
}
int64_t bscnav_vit_float_elems(const bscnav_vit_cfg_t* c) {
  VitLayout L;
  if (vit_layout(c, &L)) return BSCNAV_EINVAL;
  return L.float_total;
}
int64_t bscnav_vit_workspace_bytes(const bscnav_vit_cfg_t* c, int32_t max_batch) {
  VitLayout L;
  if (vit_layout(c, &L) || max_batch <= 0) return BSCNAV_EINVAL;
  Workspace w;
  workspace_layout(c, max_batch, &w, nullptr);
  return w.total;
}

int bscnav_vit_create(const bscnav_vit_cfg_t* c, const void* w_half, const float* w_float, void* workspace,
                      int64_t workspace_bytes, int32_t max_batch, void** out_engine) {
  if (!c || !w_half || !w_float || !workspace || !out_engine || max_batch <= 0) return BSCNAV_EINVAL;
  VitEngine* e = new VitEngine();
  e->cfg = *c;
  if (int r = vit_layout(c, &e->L)) { delete e; return r; }
  workspace_layout(c, max_batch, &e->ws, static_cast<char*>(workspace));
  if (e->ws.total > workspace_bytes) { delete e; return BSCNAV_EINVAL; }
  e->wh = static_cast<const __half*>(w_half);
  e->wf = w_float;
  e->max_batch = max_batch;
  const bool text = c->text_context > 0;
  const int G = text ? 0 : c->image / c->patch;
  e->Np = G * G;
  e->T = text ? c->text_context : 1 + c->n_reg + e->Np;
  const long long D = c->dim, M = c->mlp, rows = (long long)max_batch * e->T;
  const long long kw = c->precision == 1 ? 2 : 1;        // operand rows are (hi | lo) pairs on the wide-range path
  int r = 0;
  if (!text) r = r ? r : make_operand_map(&e->map_a0, e->ws.a0, (long long)max_batch * e->Np, kw * c->patch_k_pad, 128);
  r = r ? r : make_operand_map(&e->map_y, e->ws.y, rows, kw * D, 128);
  r = r ? r : make_operand_map(&e->map_ao, e->ws.ao, rows, kw * D, 128);
  r = r ? r : make_operand_map(&e->map_h, e->ws.hmid, rows, kw * M, 128);
  // default: the warp-specialised tcgen05 kernel (70.6 us vs 82.0 us for mma.sync at 55 frames x 16 heads, T = 257);
  // BSCNAV_ATTENTION=0|1|2 selects mma.sync / tcgen05 v1 / tcgen05 warp-specialised (tests exercise all three)
  const char* impl_env = getenv("BSCNAV_ATTENTION");
  e->attn_impl = impl_env ? atoi(impl_env) : 3;
  if (c->head_dim && c->head_dim != 64) e->attn_impl = 0;            // the tcgen05 kernels: head dim 64
  if (text && e->attn_impl != 3) e->attn_impl = 0;                   // causal mask: the second-generation kernel or mma.sync
  if (c->precision == 1) e->attn_impl = 0;                           // wide-range path: fp32 attention (run_layers_wide)
  if (e->attn_impl == 3 && !attention_ws2_supported(e->T)) e->attn_impl = 2;
  if (e->attn_impl == 2 && !attention_ws_supported(e->T)) e->attn_impl = 0;
  if (e->attn_impl == 1 && !attention_tc_supported(e->T)) e->attn_impl = 0;
  if (c->precision == 0) {
    r = r ? r : make_qkv_map3d(&e->qkv3d128, e->ws.qkv, max_batch, e->T, 3 * D, 128);
    r = r ? r : make_qkv_map3d(&e->qkv3d16, e->ws.qkv, max_batch, e->T, 3 * D, 16);
    r = r ? r : make_out_map3d(&e->ao3d, e->ws.ao, max_batch, e->T, D);
    r = r ? r : make_operand_map(&e->qkv128, e->ws.qkv, rows, 3 * D, 128);
    r = r ? r : make_operand_map(&e->qkv16, e->ws.qkv, rows, 3 * D, 16);
    r = r ? r : make_store_map(&e->out_qkv, e->ws.qkv, rows, 3 * D, false);
    r = r ? r : make_store_map(&e->out_h, e->ws.hmid, rows, M, false);
  }
  r = r ? r : make_store_map(&e->out_x, e->ws.x, rows, D, true);
  if (!text) r = r ? r : make_operand_map(&e->map_patch_w, e->wh + e->L.patch_w, D, kw * c->patch_k_pad, gemm_block_n((int)D));
  if (!text && gemm_block_n((int)D) == 256) {
    r = r ? r : make_operand_map(&e->half_patch_w, e->wh + e->L.patch_w, D, kw * c->patch_k_pad, 128);
    e->has_half_patch = true;
  }
  for (int i = 0; i < c->depth && !r; ++i) {
    const LayerOff& o = e->L.layer[i];
    r = r ? r : make_operand_map(&e->map_qkv[i], e->wh + o.qkv_w, 3 * D, kw * D, gemm_block_n((int)(3 * D)));
    r = r ? r : make_operand_map(&e->map_proj[i], e->wh + o.proj_w, D, kw * D, gemm_block_n((int)D));
    r = r ? r : make_operand_map(&e->map_fc1[i], e->wh + o.fc1_w, M, kw * D, gemm_block_n((int)M));
    r = r ? r : make_operand_map(&e->map_fc2[i], e->wh + o.fc2_w, D, kw * M, gemm_block_n((int)D));
    e->has_half[0] = gemm_block_n((int)(3 * D)) == 256; e->has_half[1] = e->has_half[3] = gemm_block_n((int)D) == 256;
    e->has_half[2] = gemm_block_n((int)M) == 256;
    if (e->has_half[0]) r = r ? r : make_operand_map(&e->half_qkv[i], e->wh + o.qkv_w, 3 * D, kw * D, 128);
    if (e->has_half[1]) r = r ? r : make_operand_map(&e->half_proj[i], e->wh + o.proj_w, D, kw * D, 128);
    if (e->has_half[2]) r = r ? r : make_operand_map(&e->half_fc1[i], e->wh + o.fc1_w, M, kw * D, 128);
    if (e->has_half[3]) r = r ? r : make_operand_map(&e->half_fc2[i], e->wh + o.fc2_w, D, kw * M, 128);
  }
  if (!r && cudaMemset(e->ws.range_flag, 0, 256) != cudaSuccess) r = BSCNAV_ECUDA;
  if (!r && cudaMemset(e->ws.kcnt, 0, kKcntLen * 4) != cudaSuccess) r = BSCNAV_ECUDA;
  if (r) { delete e; return r; }
  *out_engine = e;
  return BSCNAV_OK;
}

void bscnav_vit_destroy(void* engine) { delete static_cast<VitEngine*>(engine); }

int bscnav_vit_range_status(void* engine, int32_t reset, int32_t* host_out, void* stream_) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !host_out) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  BSC_CUDA(cudaMemcpyAsync(host_out, e->ws.range_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  if (reset) BSC_CUDA(cudaMemsetAsync(e->ws.range_flag, 0, sizeof(int32_t), stream));
  BSC_CUDA(cudaStreamSynchronize(stream));
  return BSCNAV_OK;
}

int bscnav_vit_profile(void* engine, int32_t enable) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e) return BSCNAV_EINVAL;
  e->profile = enable != 0;
  return BSCNAV_OK;
}

int bscnav_vit_profile_read(void* engine, double* out_ms, double* out_work, int32_t* out_launches) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !out_ms || !out_work || !out_launches) return BSCNAV_EINVAL;
  for (int i = 0; i < kCatCount; ++i) { out_ms[i] = 0; out_work[i] = 0; out_launches[i] = 0; }
  for (auto& sp : e->spans) {
    BSC_CUDA(cudaEventSynchronize(sp.e1));
    float ms = 0.f;
    BSC_CUDA(cudaEventElapsedTime(&ms, sp.e0, sp.e1));
    out_ms[sp.cat] += ms; out_work[sp.cat] += sp.work; out_launches[sp.cat] += 1;
    cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1);
  }
  e->spans.clear();
  return BSCNAV_OK;
}

static int clear_a0_once(VitEngine* e, cudaStream_t stream) {
  if (e->cfg.text_context > 0) return BSCNAV_EINVAL;     // text engine: image entry points do not apply
  // the K padding columns of the patch rows are never written by im2col: zero them once
  if (!e->a0_cleared) {
    BSC_CUDA(cudaMemsetAsync(e->ws.a0, 0, (size_t)(e->cfg.precision == 1 ? 2 : 1) * e->max_batch * e->Np * e->cfg.patch_k_pad * 2, stream));
    e->a0_cleared = true;
  }
  return BSCNAV_OK;
}

static int im2col_any(VitEngine* e, const float* images, int batch, cudaStream_t stream) {
  if (e->cfg.precision == 1) return im2col_images_split(images, batch, e->cfg.image, e->cfg.patch, e->cfg.patch_k_pad, e->ws.a0, stream);
  return im2col_images(images, batch, e->cfg.image, e->cfg.patch, e->cfg.patch_k_pad, e->ws.a0, stream);
}

int bscnav_vit_forward_frames(void* engine, const uint8_t* rgb, int32_t batch, int32_t height, int32_t width,
                              int32_t channels, float* out_tokens, float* out_pre, void* stream_) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !rgb || !out_tokens || batch <= 0 || batch > e->max_batch) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int r = clear_a0_once(e, stream)) return r;
  if (e->cfg.precision == 1) {
    // wide-range path: the fp32 pre-processed frames are kept and split into (hi | lo) patch rows
    { ProfScope ps(e, stream, kCatPre, (double)batch * height * width * channels);
      if (int r = preprocess_frames(rgb, batch, height, width, channels, e->cfg.image, e->cfg.patch, e->cfg.patch_k_pad, e->ws.a0h,
                                    e->ws.pre_img, e->ws.pre_tab, stream)) return r; }
    if (out_pre) BSC_CUDA(cudaMemcpyAsync(out_pre, e->ws.pre_img, (size_t)batch * 3 * e->cfg.image * e->cfg.image * 4, cudaMemcpyDeviceToDevice, stream));
    if (int r = im2col_images_split(e->ws.pre_img, batch, e->cfg.image, e->cfg.patch, e->cfg.patch_k_pad, e->ws.a0, stream)) return r;
    return vit_run(e, batch, out_tokens, stream);
  }
  { ProfScope ps(e, stream, kCatPre, (double)batch * height * width * channels);
    if (int r = preprocess_frames(rgb, batch, height, width, channels, e->cfg.image, e->cfg.patch, e->cfg.patch_k_pad, e->ws.a0,
                                  out_pre, e->ws.pre_tab, stream)) return r; }
  return vit_run(e, batch, out_tokens, stream);
}

int bscnav_vit_forward_images(void* engine, const float* images, int32_t batch, float* out_tokens, void* stream_) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !images || !out_tokens || batch <= 0 || batch > e->max_batch) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int r = clear_a0_once(e, stream)) return r;
  if (int r = im2col_any(e, images, batch, stream)) return r;
  return vit_run(e, batch, out_tokens, stream);
}

int bscnav_vit_encode_images(void* engine, const float* images, int32_t batch, const float* proj, int32_t embed_dim,
                             float* out_embed, void* stream_) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !images || !proj || !out_embed || embed_dim <= 0 || batch <= 0 || batch > e->max_batch) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int r = clear_a0_once(e, stream)) return r;
  if (int r = im2col_any(e, images, batch, stream)) return r;
  return vit_run(e, batch, nullptr, stream, proj, embed_dim, out_embed);
}

int bscnav_text_encode(void* engine, const int32_t* token_ids, int32_t batch, const float* token_embedding, int32_t vocab,
                       const float* proj, int32_t embed_dim, float* out_embed, void* stream_) {
  VitEngine* e = static_cast<VitEngine*>(engine);
  if (!e || !token_ids || !token_embedding || !proj || !out_embed || vocab <= 0 || embed_dim <= 0) return BSCNAV_EINVAL;
  if (e->cfg.text_context <= 0 || batch <= 0 || batch > e->max_batch) return BSCNAV_EINVAL;
  cudaStream_t stream = (cudaStream_t)stream_;
  const bscnav_vit_cfg_t& c = e->cfg;
  if (int r = text_embed(token_ids, token_embedding, vocab, e->wf + e->L.pos, batch, e->T, c.dim, e->ws.x, stream)) return r;
  if (int r = run_blocks(e, batch, stream)) return r;
  return cls_embed(e->ws.x, e->wf + e->L.ln_post_w, e->wf + e->L.ln_post_b, c.ln_eps, batch, e->T, c.dim, proj, embed_dim, out_embed,
                   token_ids, stream);
}

}  // extern "C"
