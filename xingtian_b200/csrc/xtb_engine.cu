// xtb_engine.cu -- C-ABI implementation (see include/xtb200.h).
#include "../../include/xtb200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "gemm_f32.cuh"
#include "optim.cuh"
#include "rl_kernels.cuh"
#include "stager.cuh"
#include "tc_dispatch.cuh"
#include <cstdlib>
#include <type_traits>

using namespace xtb;

// ------------------------------------------------------------------------------------------
// errors / bookkeeping
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess)                                                                   \
      return fail(XTB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
  } while (0)
#define LAUNCH_CHECK()                                                                       \
  do {                                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t e_ = cudaPeekAtLastError();                                                  \
    if (e_ != cudaSuccess)                                                                   \
      return fail(XTB_ERR_CUDA, "%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
  } while (0)

extern "C" int xtb_version(void) { return XTB_VERSION; }
extern "C" const char* xtb_last_error(void) { return g_err.c_str(); }
extern "C" long long xtb_launch_count(void) { return g_launches.load(); }

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// network
// ------------------------------------------------------------------------------------------
struct LayerPlan {
  xtb_layer_desc d;
  ConvGeom g{};          // conv only
  int in_size = 0;       // floats per sample of the source tensor
  int out_size = 0;
  int K = 0, N = 0;      // weight matrix [K,N]
  long long w_off = 0, b_off = 0;
  int src_act = 0;       // activation of the producing layer of the source tensor
  // device tables (conv)
  int* koff = nullptr; int* kyx = nullptr;            // forward / wgrad, indexed by k=(ky,kx,ci)
  int* dkyx = nullptr; int* dco = nullptr; int* wk = nullptr;  // dgrad, indexed by k=(ky,kx,co)
  int Kd = 0;            // KH*KW*Cout
  int sshift = 0;
  int cshift = -1, oshift = -1;   // log2(Cin), log2(Cout) when powers of two (tensor-core gathers), else -1
  bool pad = false;
  // space-to-depth view of a stride-4, 4-channel first layer (see tc::decode_u8_s2d_kernel): geometry of the equivalent
  // (k/4 x k/4, stride 1) conv over [H/4, W/4, 64]; its weight rows map to the real ones through tc::s2d_row
  bool s2d = false;
  ConvGeom g2{};
};

struct xtb_net {
  xtb_net_desc desc;
  int max_batch = 0;
  std::vector<LayerPlan> L;
  std::vector<int> tsize;       // per tensor floats/sample (0 = obs)
  long long n_params = 0;
  size_t ws_bytes = 0;
  std::vector<size_t> out_off, gout_off;  // byte offsets in workspace per tensor (fp32)
  std::vector<size_t> ohi_off, ghi_off;   // byte offsets of the bf16 hi planes (lo plane follows at plane_elems)
  std::vector<long long> plane_elems;     // elements per bf16 plane of tensor t (max_batch*tsize rounded to 8)
  size_t obs_hi_off = 0, w_hi_off = 0;    // decoded-frame plane, weight planes
  long long w_plane_elems = 0;
  bool s2d = false;                       // the observation plane is stored space-to-depth (every first-layer conv agrees)
  bool gout_planes_fresh = false;         // set while a caller guarantees head-gradient planes are current
  float* params = nullptr; float* grads = nullptr; char* ws = nullptr;
};

static int same_pad(int size, int k, int s, int* out, int* before) {
  int o = (size + s - 1) / s;
  int total = (o - 1) * s + k - size;
  if (total < 0) total = 0;
  *out = o; *before = total / 2;
  return 0;
}

extern "C" int xtb_net_create(const xtb_net_desc* desc, int max_batch, xtb_net** out) {
  if (!desc || !out || max_batch <= 0) return fail(XTB_ERR_ARG, "xtb_net_create: null/invalid argument");
  if (desc->n_layers <= 0 || desc->n_layers > XTB_MAX_LAYERS) return fail(XTB_ERR_ARG, "n_layers out of range");
  auto* net = new xtb_net();
  net->desc = *desc;
  net->max_batch = max_batch;
  struct Shape { int h, w, c; };
  std::vector<Shape> shp(desc->n_layers + 1);
  std::vector<int> tact(desc->n_layers + 1, 0);
  shp[0] = {desc->in_h, desc->in_w, desc->in_c};
  net->tsize.resize(desc->n_layers + 1);
  net->tsize[0] = desc->in_h * desc->in_w * desc->in_c;
  long long off = 0;
  for (int i = 0; i < desc->n_layers; i++) {
    LayerPlan lp;
    lp.d = desc->layers[i];
    const auto& d = lp.d;
    if (d.src < 0 || d.src > i) { delete net; return fail(XTB_ERR_ARG, "layer %d: bad src %d", i, d.src); }
    Shape is = shp[d.src];
    lp.in_size = is.h * is.w * is.c;
    lp.src_act = tact[d.src];
    if (d.kind == XTB_CONV && !d.pad_same && d.k == is.h && d.k == is.w) {
      // a VALID conv whose window covers the whole map (ImpalaCnnOpt's 11x11) is a dense layer on the
      // HWC-flattened input with the identical [kh*kw*cin, cout] weight matrix
      lp.d.kind = XTB_DENSE;
    }
    if (d.kind == XTB_CONV) {
      if (d.stride != 1 && d.stride != 2 && d.stride != 4) { delete net; return fail(XTB_ERR_ARG, "layer %d: stride must be 1,2,4", i); }
      ConvGeom& g = lp.g;
      g.H = is.h; g.W = is.w; g.C = is.c; g.KH = g.KW = d.k; g.S = d.stride; g.Cout = d.cout;
      if (d.pad_same) {
        same_pad(g.H, d.k, d.stride, &g.OH, &g.padT);
        same_pad(g.W, d.k, d.stride, &g.OW, &g.padL);
        lp.pad = true;
      } else {
        g.OH = (g.H - d.k) / d.stride + 1; g.OW = (g.W - d.k) / d.stride + 1; g.padT = g.padL = 0;
      }
      if (g.OH <= 0 || g.OW <= 0) { delete net; return fail(XTB_ERR_ARG, "layer %d: empty conv output", i); }
      g.K = d.k * d.k * g.C; g.P = g.OH * g.OW;
      g.mP = fastdiv_magic(g.P); g.mOW = fastdiv_magic(g.OW); g.mHW = fastdiv_magic(g.H * g.W); g.mW = fastdiv_magic(g.W);
      {
        // exact range of the fast division used by the tensor-core gathers: numerator < 2^24, divisor < 2^16
        long long rows = (long long)max_batch * std::max(g.P, d.src != 0 ? g.H * g.W : 0);
        if (rows >= (1LL << 24) || g.P >= (1 << 16) || g.H * g.W >= (1 << 16)) {
          delete net;
          return fail(XTB_ERR_ARG, "layer %d: max_batch*positions (%lld) exceeds 2^24 -- lower max_batch", i, rows);
        }
      }
      lp.K = g.K; lp.N = d.cout; lp.Kd = d.k * d.k * d.cout;
      lp.sshift = d.stride == 1 ? 0 : (d.stride == 2 ? 1 : 2);
      for (int sft = 0; sft < 12; sft++) { if ((1 << sft) == g.C) lp.cshift = sft; if ((1 << sft) == d.cout) lp.oshift = sft; }
      shp[i + 1] = {g.OH, g.OW, d.cout};
      // tables
      std::vector<int> koff(g.K), kyx(g.K), dkyx(lp.Kd), dco(lp.Kd), wk(lp.Kd);
      for (int ky = 0; ky < d.k; ky++)
        for (int kx = 0; kx < d.k; kx++) {
          for (int ci = 0; ci < g.C; ci++) {
            int k = (ky * d.k + kx) * g.C + ci;
            koff[k] = (ky * g.W + kx) * g.C + ci;
            kyx[k] = pack_yx(ky, kx);
          }
          for (int co = 0; co < d.cout; co++) {
            int k = (ky * d.k + kx) * d.cout + co;
            dkyx[k] = pack_yx(ky, kx);
            dco[k] = co;
            wk[k] = (ky * d.k + kx) * g.C * d.cout + co;
          }
        }
      auto up = [&](int** dst, const std::vector<int>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, v.size() * sizeof(int));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice);
      };
      cudaError_t e;
      if ((e = up(&lp.koff, koff)) != cudaSuccess || (e = up(&lp.kyx, kyx)) != cudaSuccess ||
          (e = up(&lp.dkyx, dkyx)) != cudaSuccess || (e = up(&lp.dco, dco)) != cudaSuccess ||
          (e = up(&lp.wk, wk)) != cudaSuccess) {
        delete net;
        return fail(XTB_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(e));
      }
      if (d.src == 0 && desc->input_u8 && !d.pad_same && g.C == 4 && d.stride == 4 && d.k % 4 == 0 && g.H % 4 == 0 && g.W % 4 == 0) {
        lp.s2d = true;
        ConvGeom& q = lp.g2;
        q = g;
        q.H = g.H / 4; q.W = g.W / 4; q.C = 64; q.KH = q.KW = d.k / 4; q.S = 1;
        q.mHW = fastdiv_magic(q.H * q.W); q.mW = fastdiv_magic(q.W);
      }
    } else if (d.kind == XTB_DENSE) {
      lp.K = lp.in_size; lp.N = d.cout;
      shp[i + 1] = {1, 1, d.cout};
    } else {
      delete net;
      return fail(XTB_ERR_ARG, "layer %d: unknown kind %d", i, d.kind);
    }
    if (lp.N <= 0) { delete net; return fail(XTB_ERR_ARG, "layer %d: zero outputs", i); }
    tact[i + 1] = d.act;
    lp.out_size = shp[i + 1].h * shp[i + 1].w * shp[i + 1].c;
    net->tsize[i + 1] = lp.out_size;
    lp.w_off = off; off += (long long)lp.K * lp.N;
    lp.b_off = off; off += lp.N;
    net->L.push_back(lp);
  }
  net->n_params = off;
  {   // one observation plane serves every first layer: space-to-depth only when all tensor-core first layers use it
    bool any = false, all = true;
    for (const auto& lp : net->L) if (lp.d.src == 0 && lp.d.kind == XTB_CONV) { any = true; all = all && lp.s2d; }
    net->s2d = any && all;
    if (!net->s2d) for (auto& lp : net->L) lp.s2d = false;
  }
  // workspace
  size_t w = 0;
  net->out_off.assign(desc->n_layers + 1, 0);
  net->gout_off.assign(desc->n_layers + 1, 0);
  net->ohi_off.assign(desc->n_layers + 1, 0);
  net->ghi_off.assign(desc->n_layers + 1, 0);
  net->plane_elems.assign(desc->n_layers + 1, 0);
  for (int t = 1; t <= desc->n_layers; t++) {
    size_t bytes = align_up((size_t)max_batch * net->tsize[t] * sizeof(float), 256);
    net->out_off[t] = w; w += bytes;
    net->gout_off[t] = w; w += bytes;
    long long pe = (long long)align_up((size_t)max_batch * net->tsize[t], 128);
    net->plane_elems[t] = pe;
    size_t pbytes = align_up((size_t)pe * 2 * sizeof(uint16_t), 256);   // hi + lo
    net->ohi_off[t] = w; w += pbytes;
    net->ghi_off[t] = w; w += pbytes;
  }
  net->obs_hi_off = w; w += align_up((size_t)max_batch * net->tsize[0] * sizeof(uint16_t), 256);
  net->w_plane_elems = (long long)align_up((size_t)net->n_params, 128);
  net->w_hi_off = w; w += align_up((size_t)net->w_plane_elems * 2 * sizeof(uint16_t), 256);
  net->ws_bytes = w;
  *out = net;
  return XTB_OK;
}

static void drop_graphs_of(const void* obj);
extern "C" void xtb_net_destroy(xtb_net* net) {
  if (!net) return;
  drop_graphs_of(net);
  for (auto& lp : net->L) {
    cudaFree(lp.koff); cudaFree(lp.kyx); cudaFree(lp.dkyx); cudaFree(lp.dco); cudaFree(lp.wk);
  }
  delete net;
}

extern "C" long long xtb_net_param_count(const xtb_net* net) { return net ? net->n_params : -1; }

extern "C" int xtb_net_layer_params(const xtb_net* net, int layer, long long* kernel_off, long long* bias_off,
                                    int* k_rows, int* n_cols) {
  if (!net || layer < 0 || layer >= (int)net->L.size()) return fail(XTB_ERR_ARG, "bad layer index");
  const auto& lp = net->L[layer];
  if (kernel_off) *kernel_off = lp.w_off;
  if (bias_off) *bias_off = lp.b_off;
  if (k_rows) *k_rows = lp.K;
  if (n_cols) *n_cols = lp.N;
  return XTB_OK;
}

extern "C" int xtb_net_tensor_size(const xtb_net* net, int t) {
  if (!net || t < 0 || t >= (int)net->tsize.size()) return -1;
  return net->tsize[t];
}

extern "C" size_t xtb_net_workspace_bytes(const xtb_net* net) { return net ? net->ws_bytes : 0; }

extern "C" int xtb_net_sync_weights(xtb_net* net, void* stream);
extern "C" int xtb_net_bind(xtb_net* net, float* params, float* grads, void* workspace, size_t workspace_bytes) {
  if (!net || !params || !workspace) return fail(XTB_ERR_ARG, "xtb_net_bind: null pointer");
  if (workspace_bytes < net->ws_bytes) return fail(XTB_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, net->ws_bytes);
  net->params = params; net->grads = grads; net->ws = (char*)workspace;
  return xtb_net_sync_weights(net, nullptr);
}

extern "C" float* xtb_net_tensor(xtb_net* net, int t) {
  if (!net || !net->ws || t < 1 || t >= (int)net->tsize.size()) return nullptr;
  return (float*)(net->ws + net->out_off[t]);
}
extern "C" float* xtb_net_tensor_grad(xtb_net* net, int t) {
  if (!net || !net->ws || t < 1 || t >= (int)net->tsize.size()) return nullptr;
  return (float*)(net->ws + net->gout_off[t]);
}

// ------------------------------------------------------------------------------------------
// GEMM dispatch
// ------------------------------------------------------------------------------------------
static const int kSMs = 148;

template <int BM, int BN, int TM, int TN, class AL, class BL, class EP>
static void launch_cfg(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int ksplit, cudaStream_t st) {
  constexpr int BK = 16;
  int kc = (K + ksplit - 1) / ksplit;
  kc = (kc + BK - 1) / BK * BK;
  int ks = (K + kc - 1) / kc;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, ks);
  XLAUNCH((gemm_f32_kernel<BM, BN, BK, TM, TN, AL, BL, EP>), grid, (BM / TM) * (BN / TN), 0, st, al, bl, ep, M, N, K, kc);
}

// split_ok: epilogue is atomic-accumulating so K may be partitioned over gridDim.z
template <class AL, class BL, class EP>
static void launch_gemm(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, bool split_ok, cudaStream_t st) {
  auto ctas = [&](int bm, int bn) { return (long long)((M + bm - 1) / bm) * ((N + bn - 1) / bn); };
  int ksplit = 1;
  if (N <= 32) {
    long long c = ctas(128, 32);
    if (c >= kSMs || split_ok) {
      if (split_ok) { ksplit = (int)((2 * kSMs + c - 1) / c); int mx = (K + 63) / 64; if (ksplit > mx) ksplit = mx; if (ksplit < 1) ksplit = 1; }
      launch_cfg<128, 32, 4, 4>(al, bl, ep, M, N, K, ksplit, st);
      return;
    }
    launch_cfg<32, 32, 2, 2>(al, bl, ep, M, N, K, 1, st);
    return;
  }
  long long c = ctas(64, 64);
  if (c >= kSMs || split_ok) {
    if (split_ok) { ksplit = (int)((2 * kSMs + c - 1) / c); int mx = (K + 63) / 64; if (ksplit > mx) ksplit = mx; if (ksplit < 1) ksplit = 1; }
    launch_cfg<64, 64, 4, 4>(al, bl, ep, M, N, K, ksplit, st);
    return;
  }
  launch_cfg<32, 32, 2, 2>(al, bl, ep, M, N, K, 1, st);
}


// ------------------------------------------------------------------------------------------
// tcgen05 path (gemm_tc.cuh).  g_tc_mode: 1 = use tensor cores where the shape is covered,
// 0 = fp32 CUDA-core kernels only (XTB_TC=0 in the environment, or xtb_set_tc_mode).
// ------------------------------------------------------------------------------------------
static int g_tc_mode = [] { const char* e = getenv("XTB_TC"); return e ? atoi(e) : 1; }();
extern "C" int xtb_set_tc_mode(int mode) { g_tc_mode = mode; return XTB_OK; }
extern "C" int xtb_get_tc_mode(void) { return g_tc_mode; }

struct Planes { tc::bf16* hi; long long lo_off; };   // lo plane = hi + lo_off (elements)
static inline Planes out_planes(const xtb_net* n, int t) { return Planes{(tc::bf16*)(n->ws + n->ohi_off[t]), n->plane_elems[t]}; }
static inline Planes gout_planes(const xtb_net* n, int t) { return Planes{(tc::bf16*)(n->ws + n->ghi_off[t]), n->plane_elems[t]}; }
static inline Planes obs_planes(const xtb_net* n) { return Planes{(tc::bf16*)(n->ws + n->obs_hi_off), 0}; }
static inline Planes w_planes(const xtb_net* n) { return Planes{(tc::bf16*)(n->ws + n->w_hi_off), n->w_plane_elems}; }

// bias + activation after a split-K forward (fp32 sums are already in `out`), planes kept current
__global__ void bias_act_kernel(float* __restrict__ out, const float* __restrict__ bias, long long total, int N,
                                float alpha, int act, tc::bf16* hi, long long lo_off) {
  pdl_wait(); pdl_trigger();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    float r = act_apply(act, alpha * out[i] + bias[i % N]);
    out[i] = r;
    tc::store_plane1(hi, lo_off, i, r);
  }
}

// first-layer convs read the decoded uint8 frames: 2 taps per 16-byte chunk need even geometry
static bool tc_conv_ok(const xtb_net* net, const LayerPlan& lp) {
  // K/8 column groups of a tile are cached in shared memory (TC_MAXCOLS); forward epilogues cannot split K
  if (!g_tc_mode || lp.d.kind != XTB_CONV || lp.N % 8 || lp.N > 64 || lp.K % 8 || lp.K / 8 > tc::TC_MAXCOLS) return false;
  if (lp.d.src == 0) {
    if (!net->desc.input_u8 || lp.cshift != 2) return false;
    return lp.g.KW % 2 == 0 && lp.g.padL % 2 == 0 && lp.g.S % 2 == 0 && lp.g.W % 2 == 0;
  }
  return lp.cshift >= 3;
}
static bool tc_dense_ok(const LayerPlan& lp) {
  return g_tc_mode && lp.d.kind == XTB_DENSE && lp.d.src != 0 && lp.K % 8 == 0 && lp.N % 8 == 0 && lp.K >= 64 && lp.N >= 16;
}
static bool tc_conv_dgrad_ok(const LayerPlan& lp) {
  if (!(g_tc_mode && lp.d.kind == XTB_CONV && lp.d.src != 0 && lp.g.C % 8 == 0 && lp.g.C <= 64 && lp.oshift >= 3)) return false;
  const bool cls = lp.g.S > 1 && lp.g.KH % lp.g.S == 0 && lp.g.KW % lp.g.S == 0;
  return (cls ? lp.Kd / (lp.g.S * lp.g.S) : lp.Kd) / 8 <= tc::TC_MAXCOLS;
}

template <bool PAD>
static cudaError_t tc_conv_fwd_t(const LayerPlan& lp, Planes x, bool x_split, Planes w, const float* b, float alpha,
                                 float* out, Planes po, int B, cudaStream_t st) {
  int M = B * lp.g.P;
  // first layer over a space-to-depth observation plane: k/4 x k/4 stride-1 conv over 64 channels, weight rows permuted
  const bool s2d = lp.s2d && !x_split;
  tc::SrcIm2col<PAD> a{x.hi, x.lo_off, s2d ? lp.g2 : lp.g, s2d ? 6 : lp.cshift, M};
  tc::SrcDense bs{w.hi + lp.w_off, w.lo_off, nullptr, lp.N, lp.K, s2d ? lp.g2.KH : 0};
  tc::TcEpiBiasAct ep{out, b, alpha, lp.d.act, lp.N, M, lp.N, po.hi, po.lo_off};
  if (x_split) {
    if (lp.N <= 16) return tc::launch_tc<16, 64, false, true, true, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
    if (lp.N <= 32) return tc::launch_tc<32, 64, false, true, true, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
    return tc::launch_tc<64, 64, false, true, true, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
  }
  if (lp.N <= 16) return tc::launch_tc<16, 64, false, true, false, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
  if (lp.N <= 32) return tc::launch_tc<32, 64, false, true, false, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
  return tc::launch_tc<64, 64, false, true, false, true>(a, bs, ep, M, lp.N, lp.K, 1, st);
}

static cudaError_t tc_dense_fwd(const LayerPlan& lp, Planes x, Planes w, const float* b, float* out, Planes po, int B,
                                cudaStream_t st) {
  tc::SrcDense a{x.hi, x.lo_off, nullptr, lp.K, B};
  tc::SrcDense bs{w.hi + lp.w_off, w.lo_off, nullptr, lp.N, lp.K};
  long long ctas = (long long)((B + 127) / 128) * ((lp.N + 63) / 64);
  int ks = lp.K >= 1024 ? tc::pick_ksplit(ctas, lp.K, 64, 148) : 1;
  if (ks > 1) {   // split-K: atomics into a zeroed buffer, then bias + activation (+ planes) in place
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)B * lp.N * sizeof(float), st);
    if (e != cudaSuccess) return e;
    tc::TcEpiAtomic ep{out, 1.f, lp.N, B, lp.N, 0};
    e = tc::launch_tc<64, 64, false, true, true, true>(a, bs, ep, B, lp.N, lp.K, ks, st);
    if (e != cudaSuccess) return e;
    long long total = (long long)B * lp.N;
    XLAUNCH(bias_act_kernel, (unsigned)((total + 255) / 256), 256, 0, st, out, b, total, lp.N, 1.f, lp.d.act, po.hi, po.lo_off);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaPeekAtLastError();
  }
  tc::TcEpiBiasAct ep{out, b, 1.f, lp.d.act, lp.N, B, lp.N, po.hi, po.lo_off};
  return tc::launch_tc<64, 64, false, true, true, true>(a, bs, ep, B, lp.N, lp.K, 1, st);
}

template <bool PAD>
static cudaError_t tc_conv_wgrad_t(const LayerPlan& lp, Planes x, bool x_split, Planes dy, float alpha, float* dw, int B,
                                   cudaStream_t st) {
  int Mr = B * lp.g.P;
  const bool s2d = lp.s2d && !x_split;
  tc::SrcIm2col<PAD> a{x.hi, x.lo_off, s2d ? lp.g2 : lp.g, s2d ? 6 : lp.cshift, Mr};
  tc::SrcDense bs{dy.hi, dy.lo_off, nullptr, lp.N, Mr};
  tc::TcEpiAtomic ep{dw, alpha, lp.N, lp.K, lp.N, s2d ? lp.g2.KH : 0};
  long long ctas = (long long)((lp.K + 127) / 128);
  int ks = tc::pick_ksplit(ctas, Mr, 64, 148);
  if (x_split) {
    if (lp.N <= 16) return tc::launch_tc<16, 64, true, true, true, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
    if (lp.N <= 32) return tc::launch_tc<32, 64, true, true, true, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
    return tc::launch_tc<64, 64, true, true, true, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
  }
  if (lp.N <= 16) return tc::launch_tc<16, 64, true, true, false, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
  if (lp.N <= 32) return tc::launch_tc<32, 64, true, true, false, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
  return tc::launch_tc<64, 64, true, true, false, true>(a, bs, ep, lp.K, lp.N, Mr, ks, st);
}

static cudaError_t tc_dense_wgrad(const LayerPlan& lp, Planes x, Planes dy, float* dw, int B, cudaStream_t st) {
  tc::SrcDense a{x.hi, x.lo_off, nullptr, lp.K, B};
  tc::SrcDense bs{dy.hi, dy.lo_off, nullptr, lp.N, B};
  tc::TcEpiAtomic ep{dw, 1.f, lp.N, lp.K, lp.N, 0};
  long long ctas = (long long)((lp.K + 127) / 128) * ((lp.N + 63) / 64);
  int ks = tc::pick_ksplit(ctas, B, 64, 148);
  return tc::launch_tc<64, 64, true, true, true, true>(a, bs, ep, lp.K, lp.N, B, ks, st);
}

static cudaError_t tc_conv_dgrad(const LayerPlan& lp, Planes dy, Planes w, float* gsrc, const float* xsrc, Planes pg, int acc,
                                 float* db_src, int B, cudaStream_t st) {
  const ConvGeom& g = lp.g;
  if (g.S > 1 && g.KH % g.S == 0 && g.KW % g.S == 0) {
    // strided transposed conv: one dense GEMM per parity class of the input position
    int nY = (g.H + g.S - 1) / g.S, nX = (g.W + g.S - 1) / g.S;
    tc::DgradClass k{fastdiv_magic(nY * nX), fastdiv_magic(nX), g.S, lp.sshift, g.padT, g.padL, g.H, g.W, nY, nX, 0, 0, 0, 0};
    int M = B * k.nY * k.nX;
    int Kc = lp.Kd / (g.S * g.S);
    tc::SrcDgradS a{dy.hi, dy.lo_off, g, lp.oshift, M, k, g.KW / g.S};
    tc::SrcWDgradS bs{w.hi + lp.w_off, w.lo_off, g.C, lp.N, lp.oshift, g.C, g.KW, g.KW / g.S, k};
    tc::TcEpiDgrad ep{gsrc, xsrc, lp.src_act, g.C, acc, M, g.C, pg.hi, pg.lo_off, db_src, 1, k};
    int cls = g.S * g.S;
    if (g.C <= 16) return tc::launch_tc<16, 64, false, false, true, true, tc::SrcDgradS, tc::SrcWDgradS, tc::TcEpiDgrad, true>(a, bs, ep, M, g.C, Kc, 1, st, cls);
    if (g.C <= 32) return tc::launch_tc<32, 64, false, false, true, true, tc::SrcDgradS, tc::SrcWDgradS, tc::TcEpiDgrad, true>(a, bs, ep, M, g.C, Kc, 1, st, cls);
    return tc::launch_tc<64, 64, false, false, true, true, tc::SrcDgradS, tc::SrcWDgradS, tc::TcEpiDgrad, true>(a, bs, ep, M, g.C, Kc, 1, st, cls);
  }
  int M = B * g.H * g.W;
  tc::SrcDgrad a{dy.hi, dy.lo_off, g, lp.oshift, lp.sshift, M};
  tc::SrcWDgrad bs{w.hi + lp.w_off, w.lo_off, g.C, lp.N, lp.oshift, g.C};
  tc::TcEpiDgrad ep{gsrc, xsrc, lp.src_act, g.C, acc, M, g.C, pg.hi, pg.lo_off, db_src, 0, tc::DgradClass{}};
  if (g.C <= 16) return tc::launch_tc<16, 64, false, false, true, true>(a, bs, ep, M, g.C, lp.Kd, 1, st);
  if (g.C <= 32) return tc::launch_tc<32, 64, false, false, true, true>(a, bs, ep, M, g.C, lp.Kd, 1, st);
  return tc::launch_tc<64, 64, false, false, true, true>(a, bs, ep, M, g.C, lp.Kd, 1, st);
}

static cudaError_t tc_dense_dgrad(const LayerPlan& lp, Planes dy, Planes w, float* gsrc, const float* xsrc, Planes pg, int acc,
                                  float* db_src, int B, cudaStream_t st) {
  tc::SrcDense a{dy.hi, dy.lo_off, nullptr, lp.N, B};
  tc::SrcDense bs{w.hi + lp.w_off, w.lo_off, nullptr, lp.N, lp.K};          // rows = kw (N' entity), cols = n (K')
  // the source of a dense layer may be a conv map [B, H*W*C]: its bias gradient is per channel, not per flattened
  // unit, so the column-sum fusion only applies when the source layer is dense itself (db_src != NULL then)
  tc::TcEpiDgrad ep{gsrc, xsrc, lp.src_act, lp.K, acc, B, lp.K, pg.hi, pg.lo_off, db_src, 0, tc::DgradClass{}};
  return tc::launch_tc<64, 64, false, false, true, true>(a, bs, ep, B, lp.K, lp.N, 1, st);
}

// keep the bf16 hi/lo planes of the bound parameters current (after an optimiser step / set_weights)
extern "C" int xtb_net_sync_weights(xtb_net* net, void* stream) {
  if (!net || !net->ws || !net->params) return fail(XTB_ERR_STATE, "xtb_net_sync_weights: net not bound");
  Planes w = w_planes(net);
  long long n = net->n_params;
  XLAUNCH(tc::split_planes_kernel, (unsigned)((n / 4 + 255) / 256 + 1), 256, 0, S(stream), net->params, w.hi, w.lo_off, n);
  LAUNCH_CHECK();
  return XTB_OK;
}

#ifdef XTB_TC_TRACE
extern "C" int xtb_tc_trace_set(unsigned long long* buf) {
  CUDA_TRY(cudaMemcpyToSymbol(tc::g_tc_trace, &buf, sizeof buf));
  return XTB_OK;
}
#endif

#ifdef XTB_TC_PROF
extern "C" int xtb_tc_prof_set(long long* buf) {
  CUDA_TRY(cudaMemcpyToSymbol(tc::g_tc_prof, &buf, sizeof buf));
  return XTB_OK;
}
#endif

// Self-test of the three operand-major combinations on plain matrices (see tests/test_gpu_tc.py):
//   mode 0: C[M,N] = A[M,K]   * B[K,N]      (A K-major, B MN-major : forward)
//   mode 1: C[M,N] = A[M,K]   * Bt[N,K]^T   (both K-major         : data gradient)
//   mode 2: C[M,N] = At[K,M]^T * B[K,N]     (both MN-major        : weight gradient)
extern "C" int xtb_tc_gemm_test(int mode, const float* a, const float* b, float* c, int M, int N, int K, int ksplit,
                                void* stream) {
  if (!a || !b || !c || M <= 0 || N <= 0 || K <= 0) return fail(XTB_ERR_ARG, "xtb_tc_gemm_test: bad argument");
  if (M % 8 || N % 8 || K % 8) return fail(XTB_ERR_ARG, "xtb_tc_gemm_test: sizes must be multiples of 8");
  cudaStream_t st = S(stream);
  long long na = (long long)M * K, nb = (long long)N * K;
  tc::bf16 *pa = nullptr, *pb = nullptr;
  CUDA_TRY(cudaMalloc(&pa, 2 * na * sizeof(tc::bf16)));
  CUDA_TRY(cudaMalloc(&pb, 2 * nb * sizeof(tc::bf16)));
  XLAUNCH(tc::split_planes_kernel, (unsigned)((na / 4 + 255) / 256 + 1), 256, 0, st, a, pa, na, na);
  XLAUNCH(tc::split_planes_kernel, (unsigned)((nb / 4 + 255) / 256 + 1), 256, 0, st, b, pb, nb, nb);
  cudaError_t e;
  tc::SrcDense as{pa, na, nullptr, mode == 2 ? M : K, mode == 2 ? K : M};
  tc::SrcDense bs{pb, nb, nullptr, mode == 1 ? K : N, mode == 1 ? N : K};
  if (ksplit > 1) {
    CUDA_TRY(cudaMemsetAsync(c, 0, (size_t)M * N * sizeof(float), st));
    tc::TcEpiAtomic ep{c, 1.f, N, M, N, 0};
    if (mode == 0) e = tc::launch_tc<64, 64, false, true, true, true>(as, bs, ep, M, N, K, ksplit, st);
    else if (mode == 1) e = tc::launch_tc<64, 64, false, false, true, true>(as, bs, ep, M, N, K, ksplit, st);
    else e = tc::launch_tc<64, 64, true, true, true, true>(as, bs, ep, M, N, K, ksplit, st);
  } else {
    tc::TcEpiStore ep{c, N, M, N};
    if (mode == 0) e = tc::launch_tc<32, 64, false, true, true, true>(as, bs, ep, M, N, K, 1, st);
    else if (mode == 1) e = tc::launch_tc<32, 64, false, false, true, true>(as, bs, ep, M, N, K, 1, st);
    else e = tc::launch_tc<32, 64, true, true, true, true>(as, bs, ep, M, N, K, 1, st);
  }
  g_launches.fetch_add(3, std::memory_order_relaxed);
  cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(pa); cudaFree(pb);
  if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "tc gemm launch: %s", cudaGetErrorString(e));
  if (e2 != cudaSuccess) return fail(XTB_ERR_CUDA, "tc gemm run: %s", cudaGetErrorString(e2));
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core layer ops
// ------------------------------------------------------------------------------------------
template <typename T>
static void conv_fwd(const LayerPlan& lp, const T* x, const int32_t* idx, const float* w, const float* b,
                     float alpha, float* out, Planes po, int B, cudaStream_t st) {
  int M = B * lp.g.P;
  BRowMajor bl{w, lp.N};
  EpiBiasAct ep{out, b, alpha, lp.d.act, lp.N, po.hi, po.lo_off};
  if (lp.pad) { AIm2col<T, true> al{x, idx, lp.g, lp.koff, lp.kyx}; launch_gemm(al, bl, ep, M, lp.N, lp.K, false, st); }
  else { AIm2col<T, false> al{x, idx, lp.g, lp.koff, lp.kyx}; launch_gemm(al, bl, ep, M, lp.N, lp.K, false, st); }
}
template <typename T>
static void dense_fwd(const LayerPlan& lp, const T* x, const int32_t* idx, const float* w, const float* b,
                      float alpha, float* out, Planes po, int B, cudaStream_t st) {
  ADense<T> al{x, idx, lp.K};
  BRowMajor bl{w, lp.N};
  EpiBiasAct ep{out, b, alpha, lp.d.act, lp.N, po.hi, po.lo_off};
  launch_gemm(al, bl, ep, B, lp.N, lp.K, false, st);
}
template <typename T>
static void conv_wgrad(const LayerPlan& lp, const T* x, const int32_t* idx, const float* dy, float alpha,
                       float* dw, int B, cudaStream_t st, bool bias_row = true) {
  int Mr = B * lp.g.P;
  BRowMajor bl{dy, lp.N};
  EpiAtomic ep{dw, alpha, lp.N};
  // rows 0..K-1 scaled by alpha (input decode scale); the bias row (K) must not be scaled:
  // it rides in the same GEMM only when alpha == 1, else colsum_kernel computes it.
  int rows = lp.K + ((alpha == 1.f && bias_row) ? 1 : 0);
  if (lp.pad) { AIm2colT<T, true> al{x, idx, lp.g, lp.koff, lp.kyx, Mr}; launch_gemm(al, bl, ep, rows, lp.N, Mr, true, st); }
  else { AIm2colT<T, false> al{x, idx, lp.g, lp.koff, lp.kyx, Mr}; launch_gemm(al, bl, ep, rows, lp.N, Mr, true, st); }
}
template <typename T>
static void dense_wgrad(const LayerPlan& lp, const T* x, const int32_t* idx, const float* dy, float alpha,
                        float* dw, int B, cudaStream_t st, bool bias_row = true) {
  ADenseT<T> al{x, idx, lp.K, lp.K};
  BRowMajor bl{dy, lp.N};
  EpiAtomic ep{dw, alpha, lp.N};
  launch_gemm(al, bl, ep, lp.K + ((alpha == 1.f && bias_row) ? 1 : 0), lp.N, B, true, st);
}

// bias gradient: db[n] = sum_m dy[m,n]
__global__ void colsum_kernel(const float* __restrict__ dy, int M, int N, float* __restrict__ db) {
  pdl_wait(); pdl_trigger();
  int n = blockIdx.x * 32 + (threadIdx.x & 31);
  int r0 = blockIdx.y * 1024 + (threadIdx.x >> 5);
  float s = 0.f;
  if (n < N)
    for (int m = r0; m < min(M, (int)(blockIdx.y + 1) * 1024); m += 8) s += dy[(long long)m * N + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int i = 0; i < 8; i++) t += red[i][threadIdx.x];
    atomicAdd(db + n, t);
  }
}

// ------------------------------------------------------------------------------------------
// per-layer operations (tensor-core kernel when the shape is covered, fp32 kernel otherwise)
// ------------------------------------------------------------------------------------------
// forward of layer i; tc_allowed = parameters are the bound ones (their bf16 planes are current)
static int op_forward(xtb_net* net, int i, const float* P, bool tc_allowed, const void* obs, const int32_t* idx, int B,
                      cudaStream_t st) {
  const LayerPlan& lp = net->L[i];
  float* out = (float*)(net->ws + net->out_off[i + 1]);
  Planes po = out_planes(net, i + 1);
  const float* w = P + lp.w_off;
  const float* b = P + lp.b_off;
  cudaError_t te = cudaSuccess;
  if (lp.d.src == 0) {
    float alpha = net->desc.scale;
    if (tc_allowed && tc_conv_ok(net, lp)) {
      // the decoded frames of this minibatch are in the observation plane (op_decode ran before)
      te = lp.pad ? tc_conv_fwd_t<true>(lp, obs_planes(net), false, w_planes(net), b, alpha, out, po, B, st)
                  : tc_conv_fwd_t<false>(lp, obs_planes(net), false, w_planes(net), b, alpha, out, po, B, st);
    } else if (net->desc.input_u8) {
      if (lp.d.kind == XTB_CONV) conv_fwd<uint8_t>(lp, (const uint8_t*)obs, idx, w, b, alpha, out, po, B, st);
      else dense_fwd<uint8_t>(lp, (const uint8_t*)obs, idx, w, b, alpha, out, po, B, st);
    } else {
      if (lp.d.kind == XTB_CONV) conv_fwd<float>(lp, (const float*)obs, idx, w, b, alpha, out, po, B, st);
      else dense_fwd<float>(lp, (const float*)obs, idx, w, b, alpha, out, po, B, st);
    }
  } else {
    const float* x = (const float*)(net->ws + net->out_off[lp.d.src]);
    Planes px = out_planes(net, lp.d.src);
    if (lp.d.kind == XTB_CONV) {
      if (tc_allowed && tc_conv_ok(net, lp))
        te = lp.pad ? tc_conv_fwd_t<true>(lp, px, true, w_planes(net), b, 1.f, out, po, B, st)
                    : tc_conv_fwd_t<false>(lp, px, true, w_planes(net), b, 1.f, out, po, B, st);
      else conv_fwd<float>(lp, x, nullptr, w, b, 1.f, out, po, B, st);
    } else {
      if (tc_allowed && tc_dense_ok(lp)) te = tc_dense_fwd(lp, px, w_planes(net), b, out, po, B, st);
      else dense_fwd<float>(lp, x, nullptr, w, b, 1.f, out, po, B, st);
    }
  }
  if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 forward launch (layer %d): %s", i, cudaGetErrorString(te));
  LAUNCH_CHECK();
  return XTB_OK;
}

// uint8 frame decode (+ minibatch gather) into the observation plane when a tensor-core first layer reads it
static int op_decode(xtb_net* net, const void* obs, const int32_t* idx, int B, cudaStream_t st) {
  bool need = false;
  for (const auto& lp : net->L) if (lp.d.src == 0 && tc_conv_ok(net, lp)) need = true;
  if (!need) return XTB_OK;
  int per = net->tsize[0];
  if (net->s2d) {
    dim3 grid((per / 16 + 255) / 256, B);
    XLAUNCH(tc::decode_u8_s2d_kernel, grid, 256, 0, st, (const uint8_t*)obs, idx, net->desc.in_h, net->desc.in_w, obs_planes(net).hi);
    LAUNCH_CHECK();
    return XTB_OK;
  }
  dim3 grid((per / 16 + 255) / 256, B);
  XLAUNCH(tc::decode_u8_kernel, grid, 256, 0, st, (const uint8_t*)obs, idx, per, obs_planes(net).hi);
  LAUNCH_CHECK();
  return XTB_OK;
}

static int op_wgrad(xtb_net* net, int i, const void* obs, const int32_t* idx, int B, cudaStream_t st, bool bias_done = false) {
  const LayerPlan& lp = net->L[i];
  int t = i + 1;
  const float* dy = (const float*)(net->ws + net->gout_off[t]);
  Planes pdy = gout_planes(net, t);
  float* dw = net->grads + lp.w_off;
  float* db = net->grads + lp.b_off;
  cudaError_t te = cudaSuccess;
  bool need_colsum = false;
  if (lp.d.src == 0) {
    float alpha = net->desc.scale;
    if (tc_conv_ok(net, lp)) {
      te = lp.pad ? tc_conv_wgrad_t<true>(lp, obs_planes(net), false, pdy, alpha, dw, B, st)
                  : tc_conv_wgrad_t<false>(lp, obs_planes(net), false, pdy, alpha, dw, B, st);
      need_colsum = true;
    } else {
      if (net->desc.input_u8) {
        if (lp.d.kind == XTB_CONV) conv_wgrad<uint8_t>(lp, (const uint8_t*)obs, idx, dy, alpha, dw, B, st, !bias_done);
        else dense_wgrad<uint8_t>(lp, (const uint8_t*)obs, idx, dy, alpha, dw, B, st, !bias_done);
      } else {
        if (lp.d.kind == XTB_CONV) conv_wgrad<float>(lp, (const float*)obs, idx, dy, alpha, dw, B, st, !bias_done);
        else dense_wgrad<float>(lp, (const float*)obs, idx, dy, alpha, dw, B, st, !bias_done);
      }
      need_colsum = alpha != 1.f;
    }
  } else {
    const float* x = (const float*)(net->ws + net->out_off[lp.d.src]);
    Planes px = out_planes(net, lp.d.src);
    if (lp.d.kind == XTB_CONV) {
      if (tc_conv_ok(net, lp)) {
        te = lp.pad ? tc_conv_wgrad_t<true>(lp, px, true, pdy, 1.f, dw, B, st) : tc_conv_wgrad_t<false>(lp, px, true, pdy, 1.f, dw, B, st);
        need_colsum = true;
      } else conv_wgrad<float>(lp, x, nullptr, dy, 1.f, dw, B, st, !bias_done);
    } else {
      if (tc_dense_ok(lp)) { te = tc_dense_wgrad(lp, px, pdy, dw, B, st); need_colsum = true; }
      else dense_wgrad<float>(lp, x, nullptr, dy, 1.f, dw, B, st, !bias_done);
    }
  }
  if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 wgrad launch (layer %d): %s", i, cudaGetErrorString(te));
  LAUNCH_CHECK();
  if (need_colsum && !bias_done) {   // bias gradient = column sums of dY
    int Mb = lp.d.kind == XTB_CONV ? B * lp.g.P : B;
    dim3 gridb((lp.N + 31) / 32, (Mb + 1023) / 1024);
    XLAUNCH(colsum_kernel, gridb, 256, 0, st, dy, Mb, lp.N, db);
    LAUNCH_CHECK();
  }
  return XTB_OK;
}

// data gradient of layer i into its source tensor (gradient wrt the source's pre-activation)
// db_src: bias gradient of the layer producing the source tensor, accumulated by the epilogue when non-NULL
static int op_dgrad(xtb_net* net, int i, int acc, int B, cudaStream_t st, float* db_src = nullptr) {
  const LayerPlan& lp = net->L[i];
  int t = i + 1, s = lp.d.src;
  const float* dy = (const float*)(net->ws + net->gout_off[t]);
  Planes pdy = gout_planes(net, t);
  const float* x = (const float*)(net->ws + net->out_off[s]);
  float* gsrc = (float*)(net->ws + net->gout_off[s]);
  Planes pg = gout_planes(net, s);
  const float* w = net->params + lp.w_off;
  cudaError_t te = cudaSuccess;
  if (lp.d.kind == XTB_CONV) {
    if (tc_conv_dgrad_ok(lp)) te = tc_conv_dgrad(lp, pdy, w_planes(net), gsrc, x, pg, acc, db_src, B, st);
    else {
      ADgrad al{dy, lp.g, lp.dkyx, lp.dco, lp.sshift};
      BConvDgrad bl{w, lp.wk, lp.N};
      EpiDgrad ep{gsrc, x, lp.src_act, lp.g.C, acc, pg.hi, pg.lo_off};
      launch_gemm(al, bl, ep, B * lp.g.H * lp.g.W, lp.g.C, lp.Kd, false, st);
    }
  } else {
    if (tc_dense_ok(lp)) te = tc_dense_dgrad(lp, pdy, w_planes(net), gsrc, x, pg, acc, db_src, B, st);
    else {
      ADense<float> al{dy, nullptr, lp.N};
      BTransposed bl{w, lp.N};
      EpiDgrad ep{gsrc, x, lp.src_act, lp.K, acc, pg.hi, pg.lo_off};
      launch_gemm(al, bl, ep, B, lp.K, lp.N, false, st);
    }
  }
  if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 dgrad launch (layer %d): %s", i, cudaGetErrorString(te));
  LAUNCH_CHECK();
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// forward / backward
// ------------------------------------------------------------------------------------------
static int net_forward_impl(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                            int batch, void* stream, unsigned skip_mask);
extern "C" int xtb_net_forward(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                               int batch, void* stream) {
  return net_forward_impl(net, params, obs, gather_idx, batch, stream, 0u);
}
static int net_forward_impl(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                            int batch, void* stream, unsigned skip_mask) {
  if (!net || !net->ws) return fail(XTB_ERR_STATE, "xtb_net_forward: net not bound");
  if (batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "batch %d out of range (max %d)", batch, net->max_batch);
  if (!obs) return fail(XTB_ERR_ARG, "obs is null");
  const float* P = params ? params : net->params;
  const bool tc_allowed = (P == net->params);   // foreign parameters have no bf16 planes: fp32 kernels
  cudaStream_t st = S(stream);
  if (tc_allowed) { int rc = op_decode(net, obs, gather_idx, batch, st); if (rc) return rc; }
  for (size_t i = 0; i < net->L.size(); i++) {
    if (skip_mask & (1u << i)) continue;
    int rc = op_forward(net, (int)i, P, tc_allowed, obs, gather_idx, batch, st);
    if (rc) return rc;
  }
  return XTB_OK;
}

static int net_backward_impl(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                             const int32_t* head_tensors, int n_heads, void* stream, unsigned skip_mask, bool zero_grads,
                             unsigned bias_done_tensors = 0u);
extern "C" int xtb_net_backward(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                                const int32_t* head_tensors, int n_heads, void* stream) {
  return net_backward_impl(net, obs, gather_idx, batch, head_tensors, n_heads, stream, 0u, true);
}
// head_tensors: tensors whose gradient buffers (fp32 AND planes) were filled by the caller / by a fp32 kernel
static int net_backward_impl(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                             const int32_t* head_tensors, int n_heads, void* stream, unsigned skip_mask, bool zero_grads,
                             unsigned bias_done_tensors) {
  if (!net || !net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_net_backward: net not bound (grads required)");
  if (batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "batch out of range");
  cudaStream_t st = S(stream);
  const int nl = (int)net->L.size();
  std::vector<char> has_grad(nl + 1, 0), written(nl + 1, 0);
  for (int h = 0; h < n_heads; h++) {
    int t = head_tensors[h];
    if (t < 1 || t > nl) return fail(XTB_ERR_ARG, "bad head tensor %d", t);
    has_grad[t] = 1; written[t] = 1;
    // gradients written from outside (loss kernels, python) only touch the fp32 buffer: refresh the planes
    // when a tensor-core kernel of this layer will read them
    const LayerPlan& lp = net->L[t - 1];
    if (!(skip_mask & (1u << (t - 1))) && ((lp.d.kind == XTB_CONV && (tc_conv_ok(net, lp) || tc_conv_dgrad_ok(lp))) || tc_dense_ok(lp)) &&
        !net->gout_planes_fresh) {
      Planes pg = gout_planes(net, t);
      long long n = (long long)batch * net->tsize[t];
      XLAUNCH(tc::split_planes_kernel, (unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st, (const float*)(net->ws + net->gout_off[t]), pg.hi, pg.lo_off, n);
      LAUNCH_CHECK();
    }
  }
  if (zero_grads) CUDA_TRY(cudaMemsetAsync(net->grads, 0, net->n_params * sizeof(float), st));
  // bias-gradient fusion: the bias gradient of the layer producing tensor s is the column sum of gout(s); when
  // every consumer of s writes its contribution through a tensor-core data-gradient epilogue, those epilogues
  // accumulate the column sums and the separate colsum launch of that layer is dropped
  std::vector<char> fuse_bias(nl + 1, 0);
  for (int s = 1; s <= nl; s++) {
    if (bias_done_tensors & (1u << s)) { fuse_bias[s] = 2; continue; }
    const LayerPlan& ps = net->L[s - 1];
    bool all = true, any = false;
    for (int j = 0; j < nl; j++) {
      if (net->L[j].d.src != s) continue;
      any = true;
      const LayerPlan& c = net->L[j];
      bool ok = !(skip_mask & (1u << j)) &&
                ((c.d.kind == XTB_CONV && tc_conv_dgrad_ok(c) && ps.d.kind == XTB_CONV && c.g.C % 16 == 0) ||
                 (c.d.kind == XTB_DENSE && tc_dense_ok(c) && ps.d.kind == XTB_DENSE && c.K % 16 == 0 && c.K <= 64));
      if (!ok) all = false;
    }
    fuse_bias[s] = (any && all) ? 1 : 0;
  }
  for (int i = nl - 1; i >= 0; i--) {
    if (skip_mask & (1u << i)) continue;
    const LayerPlan& lp = net->L[i];
    int t = i + 1;
    if (!has_grad[t]) continue;   // tensor does not influence the loss
    int rc = op_wgrad(net, i, obs, gather_idx, batch, st, fuse_bias[t] != 0);
    if (rc) return rc;
    if (lp.d.src != 0) {
      int s = lp.d.src;
      float* db_src = fuse_bias[s] == 1 ? net->grads + net->L[s - 1].b_off : nullptr;
      rc = op_dgrad(net, i, written[s] ? 1 : 0, batch, st, db_src);
      if (rc) return rc;
      written[s] = 1; has_grad[s] = 1;
    }
  }
  return XTB_OK;
}

// Launch ONE kernel of one layer (0 = forward, 1 = weight gradient, 2 = data gradient) on the tensors
// currently in the workspace: lets bench.py time the dominant kernel alone with CUDA events.
extern "C" int xtb_net_bench_layer(xtb_net* net, int layer, int which, const void* obs, const int32_t* gather_idx,
                                   int batch, void* stream) {
  if (!net || !net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_net_bench_layer: net not bound");
  if (layer < 0 || layer >= (int)net->L.size() || batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "bad layer/batch");
  cudaStream_t st = S(stream);
  if (which == 0) return op_forward(net, layer, net->params, true, obs, gather_idx, batch, st);
  if (which == 1) return op_wgrad(net, layer, obs, gather_idx, batch, st);
  if (which == 2) {
    if (net->L[layer].d.src == 0) return fail(XTB_ERR_ARG, "layer reads the observation: no data gradient");
    return op_dgrad(net, layer, 0, batch, st);
  }
  if (which == 3) return op_decode(net, obs, gather_idx, batch, st);
  return fail(XTB_ERR_ARG, "xtb_net_bench_layer: which must be 0..3");
}

// ------------------------------------------------------------------------------------------
// heads, GAE, losses
// ------------------------------------------------------------------------------------------
extern "C" int xtb_categorical_sample(const float* logits, int batch, int adim, const float* uniforms,
                                      uint64_t seed, uint64_t offset, int32_t* action, float* logp, void* stream) {
  if (!logits || !action || !logp || batch <= 0 || adim <= 0) return fail(XTB_ERR_ARG, "xtb_categorical_sample: bad argument");
  XLAUNCH(sample_kernel, (batch + 127) / 128, 128, 0, S(stream), logits, batch, adim, uniforms, seed, offset, action, logp);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_argmax(const float* q, int batch, int adim, int32_t* action, void* stream) {
  if (!q || !action || batch <= 0 || adim <= 0) return fail(XTB_ERR_ARG, "xtb_argmax: bad argument");
  XLAUNCH(argmax_kernel, (batch + 127) / 128, 128, 0, S(stream), q, batch, adim, action);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_gae(const float* value, const float* reward, const uint8_t* done, int n_env, int n_step,
                       float gamma, float lam, int sign_clip, float* adv, float* old_value, float* target_value,
                       void* stream) {
  if (!value || !reward || !done || !adv || !old_value || !target_value) return fail(XTB_ERR_ARG, "xtb_gae: null pointer");
  if (n_env == 0 || n_step == 0) return XTB_OK;   // empty rollout: nothing to do
  if (n_env < 0 || n_step < 0) return fail(XTB_ERR_ARG, "xtb_gae: negative size");
  int threads = 128;  // 4 envs per block
  int blocks = (n_env * 32 + threads - 1) / threads;
  XLAUNCH(gae_kernel, blocks, threads, 0, S(stream), value, reward, done, n_env, n_step, gamma, lam, sign_clip, adv,
                                                old_value, target_value);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_ppo_loss_grad(const float* logits, const float* v, const int32_t* gather_idx,
                                 const int32_t* action, const float* old_logp, const float* adv,
                                 const float* old_v, const float* target_v, int batch, int adim,
                                 const xtb_ppo_hyper* hp, float inv_count, float* dlogits, float* dv,
                                 float* loss_out, void* stream) {
  if (!logits || !v || !action || !old_logp || !adv || !old_v || !target_v || !hp || !dlogits || !dv || !loss_out)
    return fail(XTB_ERR_ARG, "xtb_ppo_loss_grad: null pointer");
  if (batch <= 0 || adim <= 0 || adim > MAX_ADIM) return fail(XTB_ERR_ARG, "xtb_ppo_loss_grad: batch/adim out of range");
  PpoHyperDev h{hp->clip_ratio, hp->ent_coef, hp->vf_clip, hp->critic_coef};
  XLAUNCH(ppo_loss_kernel, (batch + 127) / 128, 128, 0, S(stream), logits, v, gather_idx, action, old_logp, adv, old_v,
                                                              target_v, batch, adim, h, inv_count, dlogits, dv, loss_out);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_vtrace_loss_grad(const float* tp_logits, const float* baseline, const float* bp_logits,
                                    const int32_t* action, const uint8_t* done, const float* reward,
                                    int n_traj, int step_len, int adim, float gamma, float* dlogits,
                                    float* dbaseline, float* vs_out, float* pg_adv_out, float* loss_out,
                                    void* stream) {
  if (!tp_logits || !baseline || !bp_logits || !action || !done || !reward || !dlogits || !dbaseline || !loss_out)
    return fail(XTB_ERR_ARG, "xtb_vtrace_loss_grad: null pointer");
  if (n_traj <= 0 || step_len < 2 || adim <= 0 || adim > MAX_ADIM) return fail(XTB_ERR_ARG, "xtb_vtrace_loss_grad: bad sizes");
  int threads = 128;
  int blocks = (n_traj * 32 + threads - 1) / threads;
  XLAUNCH(vtrace_kernel, blocks, threads, 0, S(stream), tp_logits, baseline, bp_logits, action, done, reward, n_traj,
                                                   step_len, adim, gamma, dlogits, dbaseline, vs_out, pg_adv_out, loss_out);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_dqn_loss_grad(const float* q, const float* q_next_target, const float* q_next_online,
                                 const int32_t* action, const float* reward, const uint8_t* done, int batch,
                                 int adim, float gamma, float inv_count, float* dq, float* y_out,
                                 float* loss_out, void* stream) {
  if (!q || !q_next_target || !action || !reward || !done || !dq || !loss_out) return fail(XTB_ERR_ARG, "xtb_dqn_loss_grad: null pointer");
  if (batch <= 0 || adim <= 0) return fail(XTB_ERR_ARG, "xtb_dqn_loss_grad: bad sizes");
  XLAUNCH(dqn_loss_kernel, (batch + 127) / 128, 128, 0, S(stream), q, q_next_target, q_next_online, action, reward, done,
                                                              batch, adim, gamma, inv_count, dq, y_out, loss_out);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_mse_loss_grad(const float* q, const float* y, int batch, int adim, float inv_count, float* dq,
                                 float* loss_out, void* stream) {
  if (!q || !y || !dq || !loss_out || batch <= 0 || adim <= 0) return fail(XTB_ERR_ARG, "xtb_mse_loss_grad: bad argument");
  long long n = (long long)batch * adim;
  XLAUNCH(mse_loss_kernel, (unsigned)((n + 127) / 128), 128, 0, S(stream), q, y, n, inv_count, dq, loss_out);
  LAUNCH_CHECK();
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// optimiser
// ------------------------------------------------------------------------------------------
struct xtb_adam {
  long long count = 0;
  float lr, beta1, beta2, eps, clip;
  int clip_mode = 0, n_seg = 0, n_blk = 0;
  float *m = nullptr, *v = nullptr;
  int* blk_seg = nullptr; long long* blk_beg = nullptr; int* blk_len = nullptr;
  double* norm_sq = nullptr; float* seg_scale = nullptr; AdamState* st = nullptr;
};

extern "C" int xtb_adam_create(long long count, float lr, float beta1, float beta2, float eps, int clip_mode,
                               float clip, const long long* seg_offsets, int n_seg, float* m, float* v,
                               xtb_adam** out) {
  if (count <= 0 || !m || !v || !out) return fail(XTB_ERR_ARG, "xtb_adam_create: bad argument");
  std::vector<long long> seg;
  if (clip_mode == XTB_CLIP_PER_TENSOR) {
    if (!seg_offsets || n_seg <= 0) return fail(XTB_ERR_ARG, "per-tensor clip needs segment offsets");
    seg.assign(seg_offsets, seg_offsets + n_seg + 1);
    if (seg.front() != 0 || seg.back() != count) return fail(XTB_ERR_ARG, "segment offsets must span [0,count]");
  } else {
    seg = {0, count};
  }
  auto* o = new xtb_adam();
  o->count = count; o->lr = lr; o->beta1 = beta1; o->beta2 = beta2; o->eps = eps; o->clip = clip;
  o->clip_mode = clip_mode; o->n_seg = (int)seg.size() - 1; o->m = m; o->v = v;
  std::vector<int> bseg, blen; std::vector<long long> bbeg;
  for (int s = 0; s < o->n_seg; s++)
    for (long long b = seg[s]; b < seg[s + 1]; b += OPT_CHUNK) {
      bseg.push_back(s); bbeg.push_back(b); blen.push_back((int)std::min<long long>(OPT_CHUNK, seg[s + 1] - b));
    }
  o->n_blk = (int)bseg.size();
  AdamState init{1.f, 1.f, 0.f, 0.f};
  cudaError_t e = cudaSuccess;
  auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  chk(cudaMalloc(&o->blk_seg, o->n_blk * sizeof(int)));
  chk(cudaMalloc(&o->blk_beg, o->n_blk * sizeof(long long)));
  chk(cudaMalloc(&o->blk_len, o->n_blk * sizeof(int)));
  chk(cudaMalloc(&o->norm_sq, o->n_seg * sizeof(double)));
  chk(cudaMalloc(&o->seg_scale, o->n_seg * sizeof(float)));
  chk(cudaMalloc(&o->st, sizeof(AdamState)));
  if (e == cudaSuccess) {
    chk(cudaMemcpy(o->blk_seg, bseg.data(), o->n_blk * sizeof(int), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(o->blk_beg, bbeg.data(), o->n_blk * sizeof(long long), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(o->blk_len, blen.data(), o->n_blk * sizeof(int), cudaMemcpyHostToDevice));
    chk(cudaMemset(o->norm_sq, 0, o->n_seg * sizeof(double)));
    chk(cudaMemcpy(o->st, &init, sizeof init, cudaMemcpyHostToDevice));
    chk(cudaMemset(m, 0, count * sizeof(float)));
    chk(cudaMemset(v, 0, count * sizeof(float)));
  }
  if (e != cudaSuccess) { xtb_adam_destroy(o); return fail(XTB_ERR_CUDA, "xtb_adam_create: %s", cudaGetErrorString(e)); }
  *out = o;
  return XTB_OK;
}

extern "C" void xtb_adam_destroy(xtb_adam* o) {
  if (!o) return;
  drop_graphs_of(o);
  cudaFree(o->blk_seg); cudaFree(o->blk_beg); cudaFree(o->blk_len);
  cudaFree(o->norm_sq); cudaFree(o->seg_scale); cudaFree(o->st);
  delete o;
}

static int adam_step_impl(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream, tc::bf16* w_hi, long long w_lo);
extern "C" int xtb_adam_step(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream) {
  return adam_step_impl(o, params, grads, grad_scale, stream, nullptr, 0);
}
// optimiser step on a network's bound parameters; the weights' bf16 planes are refreshed by the same kernel
extern "C" int xtb_adam_step_net(xtb_adam* o, xtb_net* net, float grad_scale, void* stream) {
  if (!net || !net->ws || !net->params || !net->grads) return fail(XTB_ERR_STATE, "xtb_adam_step_net: net not bound");
  if (!o || o->count != net->n_params) return fail(XTB_ERR_ARG, "xtb_adam_step_net: optimiser/net size mismatch");
  Planes w = w_planes(net);
  return adam_step_impl(o, net->params, net->grads, grad_scale, stream, w.hi, w.lo_off);
}
static int adam_step_impl(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream, tc::bf16* w_hi, long long w_lo) {
  if (!o || !params || !grads) return fail(XTB_ERR_ARG, "xtb_adam_step: null pointer");
  cudaStream_t st = S(stream);
  XLAUNCH(sqnorm_kernel, o->n_blk, OPT_THREADS, 0, st, grads, o->blk_seg, o->blk_beg, o->blk_len, o->norm_sq);
  LAUNCH_CHECK();
  XLAUNCH(adam_prep_kernel, 1, 32, 0, st, o->st, o->norm_sq, o->seg_scale, o->n_seg, o->clip_mode, o->clip, grad_scale,
                                     o->lr, o->beta1, o->beta2);
  LAUNCH_CHECK();
  XLAUNCH(adam_kernel, o->n_blk, OPT_THREADS, 0, st, params, grads, o->m, o->v, o->blk_seg, o->blk_beg, o->blk_len,
                                                o->seg_scale, o->st, o->beta1, o->beta2, o->eps, w_hi, w_lo);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" const float* xtb_adam_grad_norm(const xtb_adam* o) { return o ? &o->st->grad_norm : nullptr; }
extern "C" int xtb_adam_set_lr(xtb_adam* o, float lr) {
  if (!o) return fail(XTB_ERR_ARG, "null optimiser");
  o->lr = lr;
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// fused PPO learner loop
// ------------------------------------------------------------------------------------------
// A captured graph bakes every kernel argument: the key carries everything that can change them.
struct GraphKey {
  const void* net; const void* opt; const void* obs; const void* perm; const void* loss; int n, b, e;
  const void* ro[5]; float hp[4]; int pi_t, v_t, fuse, tc; const void* ws;
  bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
};
struct GraphVal { cudaGraphExec_t exec; long long kernels; };
static std::map<GraphKey, GraphVal> g_graphs;

static int g_fuse_heads = [] { const char* e = getenv("XTB_FUSE_HEADS"); return e ? atoi(e) : 1; }();
extern "C" int xtb_set_fuse_heads(int on) { g_fuse_heads = on; return XTB_OK; }
static xtb_grad_hook g_grad_hook = nullptr;
static void* g_grad_hook_user = nullptr;
extern "C" int xtb_set_grad_hook(xtb_grad_hook hook, void* user) {
  g_grad_hook = hook; g_grad_hook_user = user;
  return XTB_OK;
}

static int ppo_train_launch(xtb_net* net, xtb_adam* opt, const xtb_ppo_rollout* ro, int N, int B, int E,
                            const int32_t* perm, const xtb_ppo_hyper* hp, int pi_t, int v_t,
                            float* loss_per_step, float inv_world, void* stream) {
  int steps_per_epoch = (N + B - 1) / B;
  CUDA_TRY(cudaMemsetAsync(loss_per_step, 0, sizeof(float) * E * steps_per_epoch, S(stream)));
  int heads[2] = {pi_t, v_t};
  int adim = net->tsize[pi_t];
  // fused heads: both heads are linear dense layers on hidden tensors of equal width
  const LayerPlan& lpi = net->L[pi_t - 1];
  const LayerPlan& lv = net->L[v_t - 1];
  int kpl = lpi.K / 32;
  bool fuse = g_fuse_heads && lpi.d.kind == XTB_DENSE && lv.d.kind == XTB_DENSE && lpi.d.act == 0 && lv.d.act == 0 &&
              lpi.d.src != 0 && lv.d.src != 0 && lpi.K == lv.K && lpi.K % 32 == 0 && adim <= 8 &&
              (kpl <= 8 || (kpl <= 16 && adim <= 4));
  unsigned skip = fuse ? ((1u << (pi_t - 1)) | (1u << (v_t - 1))) : 0u;
  int step = 0;
  for (int e = 0; e < E; e++) {
    for (int s0 = 0; s0 < N; s0 += B, step++) {
      int mb = std::min(B, N - s0);
      const int32_t* idx = perm + (long long)e * N + s0;
      int rc = net_forward_impl(net, nullptr, ro->obs, idx, mb, stream, skip);
      if (rc) return rc;
      if (fuse) {
        CUDA_TRY(cudaMemsetAsync(net->grads, 0, net->n_params * sizeof(float), S(stream)));
        PpoHeadsArgs a;
        a.h_pi = (const float*)(net->ws + net->out_off[lpi.d.src]); a.h_v = (const float*)(net->ws + net->out_off[lv.d.src]);
        a.g_pi = (float*)(net->ws + net->gout_off[lpi.d.src]); a.g_v = (float*)(net->ws + net->gout_off[lv.d.src]);
        a.gp_hi = gout_planes(net, lpi.d.src).hi; a.gp_lo = gout_planes(net, lpi.d.src).lo_off;
        a.gv_hi = gout_planes(net, lv.d.src).hi; a.gv_lo = gout_planes(net, lv.d.src).lo_off;
        a.w_pi = net->params + lpi.w_off; a.b_pi = net->params + lpi.b_off; a.w_v = net->params + lv.w_off; a.b_v = net->params + lv.b_off;
        a.dw_pi = net->grads + lpi.w_off; a.db_pi = net->grads + lpi.b_off; a.dw_v = net->grads + lv.w_off; a.db_v = net->grads + lv.b_off;
        // the hidden layers' bias gradients (column sums of g) when they are dense and only feed the heads
        auto only_feeds_heads = [&](int tsr) { for (int j = 0; j < (int)net->L.size(); j++) if (net->L[j].d.src == tsr && !(skip & (1u << j))) return false; return true; };
        bool bh_pi_ok = net->L[lpi.d.src - 1].d.kind == XTB_DENSE && only_feeds_heads(lpi.d.src);
        bool bh_v_ok = net->L[lv.d.src - 1].d.kind == XTB_DENSE && only_feeds_heads(lv.d.src);
        a.dbh_pi = bh_pi_ok ? net->grads + net->L[lpi.d.src - 1].b_off : nullptr;
        a.dbh_v = bh_v_ok ? net->grads + net->L[lv.d.src - 1].b_off : nullptr;
        unsigned bias_done = (bh_pi_ok ? (1u << lpi.d.src) : 0u) | ((lpi.d.src != lv.d.src && bh_v_ok) ? (1u << lv.d.src) : 0u);
        a.idx = idx; a.action = ro->action; a.old_logp = ro->old_logp; a.adv = ro->adv; a.old_v = ro->old_v; a.target_v = ro->target_v;
        a.logits_out = xtb_net_tensor(net, pi_t); a.v_out = xtb_net_tensor(net, v_t); a.loss_out = loss_per_step + step;
        a.B = mb; a.K = lpi.K; a.A = adim; a.act_pi = lpi.src_act; a.act_v = lv.src_act; a.shared = lpi.d.src == lv.d.src ? 1 : 0;
        a.hp = PpoHyperDev{hp->clip_ratio, hp->ent_coef, hp->vf_clip, hp->critic_coef}; a.inv_count = inv_world / mb;
        int blocks = std::max(1, std::min(148, (mb + 15) / 16));
        size_t shb = (size_t)(lpi.K * (adim + 1) + adim + 1) * sizeof(float);
        if (kpl <= 2) XLAUNCH((ppo_heads_kernel<2, 8>), blocks, 256, shb, S(stream), a);
        else if (kpl <= 8) XLAUNCH((ppo_heads_kernel<8, 8>), blocks, 256, shb, S(stream), a);
        else XLAUNCH((ppo_heads_kernel<16, 4>), blocks, 256, shb, S(stream), a);
        LAUNCH_CHECK();
        int srcs[2] = {lpi.d.src, lv.d.src};
        net->gout_planes_fresh = true;    // the fused kernel wrote fp32 and planes
        rc = net_backward_impl(net, ro->obs, idx, mb, srcs, a.shared ? 1 : 2, stream, skip, false, bias_done);
        net->gout_planes_fresh = false;
        if (rc) return rc;
      } else {
        rc = xtb_ppo_loss_grad(xtb_net_tensor(net, pi_t), xtb_net_tensor(net, v_t), idx, ro->action, ro->old_logp,
                               ro->adv, ro->old_v, ro->target_v, mb, adim, hp, inv_world / mb,
                               xtb_net_tensor_grad(net, pi_t), xtb_net_tensor_grad(net, v_t), loss_per_step + step, stream);
        if (rc) return rc;
        rc = xtb_net_backward(net, ro->obs, idx, mb, heads, 2, stream);
        if (rc) return rc;
      }
      if (g_grad_hook) {
        rc = g_grad_hook(g_grad_hook_user, net->grads, net->n_params, stream);
        if (rc) return fail(XTB_ERR_STATE, "gradient hook failed with %d", rc);
      }
      rc = xtb_adam_step_net(opt, net, 1.f, stream);
      if (rc) return rc;
    }
  }
  return XTB_OK;
}

extern "C" int xtb_ppo_train(xtb_net* net, xtb_adam* opt, const xtb_ppo_rollout* ro, int n_sample,
                             int batch_size, int n_epoch, const int32_t* perm, const xtb_ppo_hyper* hp,
                             int pi_tensor, int v_tensor, float* loss_per_step, int use_graph, void* stream) {
  if (!net || !opt || !ro || !perm || !hp || !loss_per_step) return fail(XTB_ERR_ARG, "xtb_ppo_train: null pointer");
  if (!net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_ppo_train: net not bound");
  if (n_sample <= 0 || batch_size <= 0 || n_epoch <= 0) return fail(XTB_ERR_ARG, "xtb_ppo_train: bad sizes");
  if (std::min(batch_size, n_sample) > net->max_batch) return fail(XTB_ERR_ARG, "batch_size exceeds net max_batch");
  int nl = (int)net->L.size();
  if (pi_tensor < 1 || pi_tensor > nl || v_tensor < 1 || v_tensor > nl || net->tsize[v_tensor] != 1)
    return fail(XTB_ERR_ARG, "xtb_ppo_train: bad head tensors");
  float inv_world = 1.f;
  if (g_grad_hook) {   // data-parallel: the hook sums gradients over ranks; every rank holds B/world samples
    int world = 1;
    int rc = g_grad_hook(g_grad_hook_user, nullptr, 0, nullptr);   // query: returns world size when grads == NULL
    if (rc > 0) world = rc;
    inv_world = 1.f / world;
  }
  if (!use_graph || !stream || g_grad_hook)
    return ppo_train_launch(net, opt, ro, n_sample, batch_size, n_epoch, perm, hp, pi_tensor, v_tensor, loss_per_step, inv_world, stream);
  GraphKey key;
  memset(&key, 0, sizeof key);
  key.net = net; key.opt = opt; key.obs = ro->obs; key.perm = perm; key.loss = loss_per_step;
  key.n = n_sample; key.b = batch_size; key.e = n_epoch;
  key.ro[0] = ro->action; key.ro[1] = ro->old_logp; key.ro[2] = ro->adv; key.ro[3] = ro->old_v; key.ro[4] = ro->target_v;
  key.hp[0] = hp->clip_ratio; key.hp[1] = hp->ent_coef; key.hp[2] = hp->vf_clip; key.hp[3] = hp->critic_coef;
  key.pi_t = pi_tensor; key.v_t = v_tensor; key.fuse = g_fuse_heads; key.tc = g_tc_mode; key.ws = net->ws;
  auto it = g_graphs.find(key);
  if (it == g_graphs.end()) {
    cudaStream_t st = S(stream);
    long long before = g_launches.load();
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = ppo_train_launch(net, opt, ro, n_sample, batch_size, n_epoch, perm, hp, pi_tensor, v_tensor, loss_per_step, 1.f, stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    long long captured = g_launches.load() - before;
    g_launches.store(before);   // captured, not launched yet
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    it = g_graphs.emplace(key, GraphVal{exec, captured}).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second.exec, S(stream)));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// rollout inference: T batched policy evaluations over the E stacked observations
// ------------------------------------------------------------------------------------------
struct InferKey {
  const void* net; const void* obs; const void* idx; const void* act; const void* logp; const void* val; const void* ctr;
  const void* ws; unsigned long long seed; int e, t, pi_t, v_t, tc;
  bool operator<(const InferKey& o) const { return memcmp(this, &o, sizeof(InferKey)) < 0; }
};
static std::map<InferKey, GraphVal> g_infer_graphs;
// cached graphs hold raw pointers into a network / optimiser: they die with the object (a later object may
// be allocated at the same address)
static void drop_graphs_of(const void* obj) {
  for (auto it = g_graphs.begin(); it != g_graphs.end();) {
    if (it->first.net == obj || it->first.opt == obj) { cudaGraphExecDestroy(it->second.exec); it = g_graphs.erase(it); } else ++it;
  }
  for (auto it = g_infer_graphs.begin(); it != g_infer_graphs.end();) {
    if (it->first.net == obj) { cudaGraphExecDestroy(it->second.exec); it = g_infer_graphs.erase(it); } else ++it;
  }
}

static int rollout_infer_launch(xtb_net* net, const void* obs, const int32_t* step_idx, int E, int T, int pi_t, int v_t,
                                uint64_t seed, unsigned long long* offset_dev, int32_t* action, float* logp, float* value,
                                void* stream) {
  int adim = net->tsize[pi_t];
  const LayerPlan& lpi = net->L[pi_t - 1];
  const LayerPlan& lv = net->L[v_t - 1];
  int kpl = lpi.K / 32;
  bool fuse = g_fuse_heads && lpi.d.kind == XTB_DENSE && lv.d.kind == XTB_DENSE && lpi.d.act == 0 && lv.d.act == 0 &&
              lpi.d.src != 0 && lv.d.src != 0 && lpi.K == lv.K && lpi.K % 32 == 0 && adim <= 8 && kpl <= 16;
  unsigned skip = fuse ? ((1u << (pi_t - 1)) | (1u << (v_t - 1))) : 0u;
  for (int t = 0; t < T; t++) {
    int rc = net_forward_impl(net, nullptr, obs, step_idx ? step_idx + (long long)t * E : nullptr, E, stream, skip);
    if (rc) return rc;
    int32_t* a_t = action + (long long)t * E; float* lp_t = logp + (long long)t * E; float* v_o = value + (long long)t * E;
    if (fuse) {
      const float* hp = (const float*)(net->ws + net->out_off[lpi.d.src]);
      const float* hv = (const float*)(net->ws + net->out_off[lv.d.src]);
      const float *wp = net->params + lpi.w_off, *bp = net->params + lpi.b_off, *wv = net->params + lv.w_off, *bv = net->params + lv.b_off;
      int blocks = std::max(1, std::min(148, (E + 7) / 8));
      if (kpl <= 2) XLAUNCH((ppo_infer_heads_kernel<2, 8>), blocks, 256, 0, S(stream), hp, hv, wp, bp, wv, bv, E, lpi.K, adim, seed, offset_dev, t, a_t, lp_t, v_o, xtb_net_tensor(net, pi_t));
      else if (kpl <= 8) XLAUNCH((ppo_infer_heads_kernel<8, 8>), blocks, 256, 0, S(stream), hp, hv, wp, bp, wv, bv, E, lpi.K, adim, seed, offset_dev, t, a_t, lp_t, v_o, xtb_net_tensor(net, pi_t));
      else XLAUNCH((ppo_infer_heads_kernel<16, 8>), blocks, 256, 0, S(stream), hp, hv, wp, bp, wv, bv, E, lpi.K, adim, seed, offset_dev, t, a_t, lp_t, v_o, xtb_net_tensor(net, pi_t));
    } else {
      XLAUNCH(sample_rollout_kernel, (E + 127) / 128, 128, 0, S(stream), xtb_net_tensor(net, pi_t), xtb_net_tensor(net, v_t), E, adim, seed,
                                                                   offset_dev, t, a_t, lp_t, v_o);
    }
    LAUNCH_CHECK();
  }
  XLAUNCH(bump_counter_kernel, 1, 1, 0, S(stream), offset_dev, T);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" int xtb_ppo_rollout_infer(xtb_net* net, const void* obs, const int32_t* step_idx, int n_env, int n_step,
                                     int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                                     int32_t* action, float* logp, float* value, int use_graph, void* stream) {
  if (!net || !net->ws || !obs || !offset_dev || !action || !logp || !value) return fail(XTB_ERR_ARG, "xtb_ppo_rollout_infer: null pointer");
  int nl = (int)net->L.size();
  if (n_env <= 0 || n_env > net->max_batch || n_step <= 0) return fail(XTB_ERR_ARG, "xtb_ppo_rollout_infer: bad sizes");
  if (pi_tensor < 1 || pi_tensor > nl || v_tensor < 1 || v_tensor > nl || net->tsize[v_tensor] != 1 || net->tsize[pi_tensor] > MAX_ADIM)
    return fail(XTB_ERR_ARG, "xtb_ppo_rollout_infer: bad head tensors");
  if (!use_graph || !stream)
    return rollout_infer_launch(net, obs, step_idx, n_env, n_step, pi_tensor, v_tensor, seed, offset_dev, action, logp, value, stream);
  InferKey key;
  memset(&key, 0, sizeof key);
  key.net = net; key.obs = obs; key.idx = step_idx; key.act = action; key.logp = logp; key.val = value; key.ctr = offset_dev;
  key.ws = net->ws; key.seed = seed; key.e = n_env; key.t = n_step; key.pi_t = pi_tensor; key.v_t = v_tensor; key.tc = g_tc_mode;
  auto it = g_infer_graphs.find(key);
  if (it == g_infer_graphs.end()) {
    cudaStream_t st = S(stream);
    long long before = g_launches.load();
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = rollout_infer_launch(net, obs, step_idx, n_env, n_step, pi_tensor, v_tensor, seed, offset_dev, action, logp, value, stream);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    long long captured = g_launches.load() - before;
    g_launches.store(before);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    it = g_infer_graphs.emplace(key, GraphVal{exec, captured}).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second.exec, S(stream)));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// staging helpers
// ------------------------------------------------------------------------------------------
extern "C" void* xtb_pinned_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { fail(XTB_ERR_NOMEM, "cudaHostAlloc(%zu) failed", bytes); return nullptr; }
  return p;
}
extern "C" void xtb_pinned_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" int xtb_copy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, S(stream)));
  return XTB_OK;
}
extern "C" int xtb_copy_h2d_staged(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes && (!dst || !src)) return fail(XTB_ERR_ARG, "xtb_copy_h2d_staged: null pointer");
  CUDA_TRY(xtb::Stager::instance().stage_h2d(dst, src, bytes, S(stream)));
  return XTB_OK;
}
// PPO.predict with host buffers in one call (xt/model/ppo/ppo.py:104-109): staged H2D of the observations, the
// (graphed) forward + sampling, one packed D2H of [action | logp | value] and a stream synchronise.
extern "C" int xtb_ppo_predict_host(xtb_net* net, const void* obs_host, size_t obs_bytes, void* obs_dev, int n_env,
                                    int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                                    float* out_dev, float* out_host, int use_graph, void* stream) {
  if (!obs_host || !obs_dev || !out_dev || !out_host) return fail(XTB_ERR_ARG, "xtb_ppo_predict_host: null pointer");
  CUDA_TRY(xtb::Stager::instance().stage_h2d(obs_dev, obs_host, obs_bytes, S(stream)));
  int rc = xtb_ppo_rollout_infer(net, obs_dev, nullptr, n_env, 1, pi_tensor, v_tensor, seed, offset_dev,
                                 reinterpret_cast<int32_t*>(out_dev), out_dev + n_env, out_dev + 2 * (size_t)n_env, use_graph, stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_host, out_dev, sizeof(float) * 3 * (size_t)n_env, cudaMemcpyDeviceToHost, S(stream)));
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  return XTB_OK;
}
extern "C" int xtb_copy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(stream)));
  return XTB_OK;
}
extern "C" int xtb_stream_sync(void* stream) {
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  return XTB_OK;
}
