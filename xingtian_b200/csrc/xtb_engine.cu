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
#include "bp_gemm.cuh"
#include "comm.cuh"
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

// ---- data-parallel communicator ---------------------------------------------------------------
static NcclApi g_nccl;
static xtb_comm* g_comm = nullptr;          // installed communicator: gradients are all-reduced inside the training loops
extern "C" int xtb_comm_unique_id(const char* nccl_path, void* id128) {
  if (!id128) return fail(XTB_ERR_ARG, "xtb_comm_unique_id: null pointer");
  if (const char* err = g_nccl.load(nccl_path)) return fail(XTB_ERR_STATE, "cannot load NCCL: %s", err);
  NcclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc) return fail(XTB_ERR_CUDA, "ncclGetUniqueId: %s", g_nccl.GetErrorString(rc));
  memcpy(id128, &id, sizeof id);
  return XTB_OK;
}
extern "C" int xtb_comm_create(const char* nccl_path, const void* id128, int rank, int world, xtb_comm** out) {
  if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return fail(XTB_ERR_ARG, "xtb_comm_create: bad argument");
  if (const char* err = g_nccl.load(nccl_path)) return fail(XTB_ERR_STATE, "cannot load NCCL: %s", err);
  auto* c = new xtb_comm();
  c->rank = rank; c->world = world;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof id);
  int rc = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (rc) { delete c; return fail(XTB_ERR_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc)); }
  cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming);
  if (e != cudaSuccess) { g_nccl.CommDestroy(c->comm); delete c; return fail(XTB_ERR_CUDA, "xtb_comm_create: %s", cudaGetErrorString(e)); }
  *out = c;
  return XTB_OK;
}
static void drop_graphs_of(const void* obj);
extern "C" void xtb_comm_destroy(xtb_comm* c) {
  if (!c) return;
  if (g_comm == c) g_comm = nullptr;
  // captured training graphs hold NCCL kernels of this communicator: they go first, and nothing may be in flight
  drop_graphs_of(c);
  cudaDeviceSynchronize();
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->fork) cudaEventDestroy(c->fork);
  if (c->join) cudaEventDestroy(c->join);
  delete c;
}
extern "C" int xtb_comm_world(const xtb_comm* c) { return c ? c->world : 1; }
// NULL uninstalls.  While installed, the fused training loops scale losses/gradients by 1/(world * B_local), sum the
// gradient bucket over ranks before the optimiser step and stay inside the CUDA graph.
extern "C" int xtb_set_grad_comm(xtb_comm* c) { g_comm = c; return XTB_OK; }
// sum `count` floats in place over the ranks of `c`, on `stream`
extern "C" int xtb_comm_allreduce(xtb_comm* c, float* buf, long long count, void* stream) {
  if (!c || !buf || count < 0) return fail(XTB_ERR_ARG, "xtb_comm_allreduce: bad argument");
  if (count == 0 || c->world == 1) return XTB_OK;
  int rc = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat, kNcclSum, c->comm, S(stream));
  if (rc) return fail(XTB_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(rc));
  return XTB_OK;
}
// Early bucket [off, off+count): forked onto the communicator's side stream at the current point of `stream`, so the
// exchange overlaps whatever is enqueued on `stream` afterwards; comm_join() makes `stream` wait for it.
static int comm_fork_allreduce(xtb_comm* c, float* buf, long long count, cudaStream_t st) {
  CUDA_TRY(cudaEventRecord(c->fork, st));
  CUDA_TRY(cudaStreamWaitEvent(c->side, c->fork, 0));
  int rc = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat, kNcclSum, c->comm, c->side);
  if (rc) return fail(XTB_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(rc));
  CUDA_TRY(cudaEventRecord(c->join, c->side));
  return XTB_OK;
}
static int comm_join(xtb_comm* c, cudaStream_t st) {
  CUDA_TRY(cudaStreamWaitEvent(st, c->join, 0));
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// network
// ------------------------------------------------------------------------------------------
static const int kSMs = 148;
static const int kMaxDynSmem = 232448 - 2048;   // 227 KiB per CTA minus the kernels' static shared memory

struct LayerPlan {
  xtb_layer_desc d;
  ConvGeom g{};          // conv only
  int in_size = 0;       // floats per sample of the source tensor
  int out_size = 0;
  int K = 0, N = 0;      // weight matrix [K,N]
  long long w_off = 0, b_off = 0;
  int src_act = 0;       // activation of the producing layer of the source tensor
  // device tables (conv, fp32 path)
  int* koff = nullptr; int* kyx = nullptr;            // forward / wgrad, indexed by k=(ky,kx,ci)
  int* dkyx = nullptr; int* dco = nullptr; int* wk = nullptr;  // dgrad, indexed by k=(ky,kx,co)
  int Kd = 0;            // KH*KW*Cout
  int sshift = 0;
  bool pad = false;
  // ---- tensor-core (batch-planar) plan; tc = the layer's shapes are covered
  bool tc = false;
  bool s2d = false;      // stride-4 4-channel first layer run as a (k/4 x k/4, stride 1) conv over a space-to-depth plane
  int k4 = 0;
  ConvGeom q{};          // geometry the tensor-core kernels use (space-to-depth view for s2d layers)
  long long blob_off = 0;   // element offset of this layer's weight blob (hi plane)
  bool w_res = false;       // conv: blob resident in shared memory
  int n_fwd = 0, n_dg = 0;  // accumulator columns: forward (Cout / dense N tile), data gradient (Cin / dense K tile)
  int run_chunks = 0, mts = 0, R = 0;   // conv weight gradient: chunks per filter row, M tiles per row, accumulators
  size_t part_off = 0, dbpart_off = 0;  // workspace offsets of the partial-sum areas
  // host-built stage walks of the tensor-core kernels (conv layers) and their workspace offsets
  std::vector<bp::StageEnt> fwd_st, dg_st; std::vector<bp::UnitEnt> fwd_un, dg_un; std::vector<bp::WgEnt> wg_tab;
  size_t fwd_st_off = 0, fwd_un_off = 0, dg_st_off = 0, dg_un_off = 0, wg_off = 0;
};

// The K-stage walk of every unit of a conv layer (see bp_rows_kernel): forward = per filter row the taps inside the
// image are one contiguous feature run, cut into stages of <= 64 elements; data gradient = one stage per filter tap
// whose output position exists.  Weight gradient: per (output pixel, accumulator) the first X chunk and which of the
// 16 chunks of the M tile are real.
static void build_conv_tables(LayerPlan& lp) {
  const ConvGeom& q = lp.q;
  const int Cout = lp.N;
  for (int oy = 0; oy < q.OH; oy++)
    for (int ox = 0; ox < q.OW; ox++) {
      uint32_t first = (uint32_t)lp.fwd_st.size(), count = 0;
      for (int ky = 0; ky < q.KH; ky++) {
        int iy = oy * q.S - q.padT + ky;
        if (iy < 0 || iy >= q.H) continue;
        int x0 = ox * q.S - q.padL, xl = std::max(x0, 0), xh = std::min(x0 + q.KW, q.W);
        if (xh <= xl) continue;
        int chunk = ((iy * q.W + xl) * q.C) >> 3, rem = ((xh - xl) * q.C) >> 3, wrow = (ky * q.KW + (xl - x0)) * q.C;
        while (rem > 0) {
          int nch = std::min(8, rem);
          lp.fwd_st.push_back(bp::StageEnt{(uint32_t)chunk, (uint16_t)wrow, (uint16_t)nch});
          chunk += nch; rem -= nch; wrow += nch * 8; count++;
        }
      }
      lp.fwd_un.push_back(first | (count << 24));
    }
  for (int iy = 0; iy < q.H; iy++)
    for (int ix = 0; ix < q.W; ix++) {
      uint32_t first = (uint32_t)lp.dg_st.size(), count = 0;
      for (int t = 0; t < q.KH * q.KW; t++) {
        int ky = t / q.KW, kx = t - ky * q.KW;
        int ty = iy + q.padT - ky, tx = ix + q.padL - kx;
        if (ty < 0 || tx < 0) continue;
        int oy = ty / q.S, ox = tx / q.S;
        if (oy * q.S != ty || ox * q.S != tx || oy >= q.OH || ox >= q.OW) continue;
        lp.dg_st.push_back(bp::StageEnt{(uint32_t)(((oy * q.OW + ox) * Cout) >> 3), (uint16_t)(t * q.C), (uint16_t)(Cout >> 3)});
        count++;
      }
      lp.dg_un.push_back(first | (count << 24));
    }
  for (int oy = 0; oy < q.OH; oy++)
    for (int ox = 0; ox < q.OW; ox++)
      for (int r = 0; r < lp.R; r++) {
        int ky = r / lp.mts, mt = r - ky * lp.mts;
        int iy = oy * q.S - q.padT + ky;
        bp::WgEnt e{0, 0, 0};
        if (iy >= 0 && iy < q.H) {
          int xc0 = ox * q.S - q.padL;
          long long chunk0 = ((long long)(iy * q.W + xc0) * q.C) / 8 + mt * 16;
          e.x_chunk = (int32_t)chunk0;
          for (int c = 0; c < 16; c++) {
            int cj = mt * 16 + c, px = xc0 + (cj * 8) / q.C;
            if (cj < lp.run_chunks && px >= 0 && px < q.W) e.okmask |= (uint16_t)(1u << c);
          }
          e.valid = 1;
        }
        lp.wg_tab.push_back(e);
      }
}

struct PendingRed { bp::RedSeg s; };

struct xtb_net {
  xtb_net_desc desc;
  int max_batch = 0, pitch = 0;
  std::vector<LayerPlan> L;
  std::vector<int> tsize;       // per tensor floats/sample (0 = obs)
  long long n_params = 0;
  size_t ws_bytes = 0;
  std::vector<size_t> out_off, gout_off;  // byte offsets in workspace per tensor (fp32 row-major)
  std::vector<size_t> obp_off, gbp_off;   // byte offsets of the batch-planar hi planes (lo plane follows at plane_elems)
  std::vector<long long> plane_elems;     // elements per plane of tensor t = tsize * pitch
  size_t obs_bp_off = 0; int obs_feats = 0;            // decoded-frame plane (space-to-depth canvas)
  int H4 = 0, W4 = 0;
  size_t blob_off = 0; long long blob_elems = 0;       // weight blobs: hi plane, lo plane follows
  size_t splitk_off = 0, zeros_off = 0, segs_off = 0, heads_part_off = 0;
  std::vector<bp::BlobSeg> blob_segs;
  bool any_tc = false;
  float* params = nullptr; float* grads = nullptr; char* ws = nullptr;
  // which representation of every tensor / tensor gradient is current
  std::vector<char> f32_ok, bp_ok, gf32_ok, gbp_ok;
  bool obs_bp_ok = false;
  std::vector<bp::RedSeg> pending;        // ordered reductions queued by the running backward pass
};

static int same_pad(int size, int k, int s, int* out, int* before) {
  int o = (size + s - 1) / s;
  int total = (o - 1) * s + k - size;
  if (total < 0) total = 0;
  *out = o; *before = total / 2;
  return 0;
}

// ------------------------------------------------------------------------------------------
// tcgen05 path (bp_gemm.cuh).  g_tc_mode: 1 = use tensor cores where the shape is covered,
// 0 = fp32 CUDA-core kernels only (XTB_TC=0 in the environment, or xtb_set_tc_mode).
// ------------------------------------------------------------------------------------------
static int g_tc_mode = [] { const char* e = getenv("XTB_TC"); return e ? atoi(e) : 1; }();
extern "C" int xtb_set_tc_mode(int mode) { g_tc_mode = mode; return XTB_OK; }
extern "C" int xtb_get_tc_mode(void) { return g_tc_mode; }

static inline bp::BpT out_bp(const xtb_net* n, int t) { return bp::BpT{(bp::bf16*)(n->ws + n->obp_off[t]), n->plane_elems[t], n->pitch}; }
static inline bp::BpT gout_bp(const xtb_net* n, int t) { return bp::BpT{(bp::bf16*)(n->ws + n->gbp_off[t]), n->plane_elems[t], n->pitch}; }
static inline bp::BpT obs_bp(const xtb_net* n) { return bp::BpT{(bp::bf16*)(n->ws + n->obs_bp_off), 0, n->pitch}; }
static inline bp::BpT no_bp() { return bp::BpT{nullptr, 0, 0}; }
static inline const bp::bf16* blob_hi(const xtb_net* n, const LayerPlan& lp) { return (const bp::bf16*)(n->ws + n->blob_off) + lp.blob_off; }
static inline bool use_tc(const LayerPlan& lp) { return g_tc_mode && lp.tc; }

static int pick_tile(int n) { return n % 64 == 0 ? 64 : (n % 32 == 0 ? 32 : (n % 16 == 0 ? 16 : 0)); }

extern "C" int xtb_net_create(const xtb_net_desc* desc, int max_batch, xtb_net** out) {
  if (!desc || !out || max_batch <= 0) return fail(XTB_ERR_ARG, "xtb_net_create: null/invalid argument");
  if (desc->n_layers <= 0 || desc->n_layers > XTB_MAX_LAYERS) return fail(XTB_ERR_ARG, "n_layers out of range");
  auto* net = new xtb_net();
  net->desc = *desc;
  net->max_batch = max_batch;
  net->pitch = (max_batch + 15) / 16 * 16;
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
    if (lp.d.kind == XTB_CONV) {
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
      lp.K = g.K; lp.N = d.cout; lp.Kd = d.k * d.k * d.cout;
      lp.sshift = d.stride == 1 ? 0 : (d.stride == 2 ? 1 : 2);
      shp[i + 1] = {g.OH, g.OW, d.cout};
      // tables (fp32 path)
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
      // ---- tensor-core plan
      lp.q = g;
      const bool cout_ok = d.cout % 16 == 0 && d.cout <= 64;
      if (d.src == 0) {
        // uint8 frames, 4 channels, stride 4: space-to-depth canvas [OH+k/4-1, OW+k/4-1] blocks of 64 features
        if (desc->input_u8 && g.C == 4 && d.stride == 4 && d.k % 4 == 0 && g.padL % 2 == 0 && g.W % 4 == 0 && cout_ok) {
          lp.tc = lp.s2d = true;
          lp.k4 = d.k / 4;
          ConvGeom& q = lp.q;
          q.H = g.OH + lp.k4 - 1; q.W = g.OW + lp.k4 - 1; q.C = 64; q.KH = q.KW = lp.k4; q.S = 1; q.padT = q.padL = 0;
        }
      } else if (g.C % 16 == 0 && g.C <= 64 && cout_ok && d.stride <= d.k) {
        lp.tc = true;
      }
      if (lp.tc) {
        const ConvGeom& q = lp.q;
        lp.n_fwd = d.cout; lp.n_dg = q.C;
        lp.run_chunks = q.KW * q.C / 8;
        lp.mts = (lp.run_chunks + 15) / 16;
        lp.R = q.KH * lp.mts;
        if (lp.R > 32 || lp.R * d.cout > 512) lp.tc = lp.s2d = false;
        lp.w_res = (long long)lp.K * lp.N * 4 <= 96 * 1024;
      }
    } else if (lp.d.kind == XTB_DENSE) {
      lp.K = lp.in_size; lp.N = d.cout;
      shp[i + 1] = {1, 1, d.cout};
      if (d.src != 0 && lp.K % 16 == 0 && pick_tile(lp.N) && pick_tile(lp.K)) {
        lp.tc = true;
        lp.n_fwd = pick_tile(lp.N); lp.n_dg = pick_tile(lp.K);
      }
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
  {   // one observation canvas serves every tensor-core first layer: they must agree on its geometry
    const LayerPlan* f = nullptr; bool agree = true;
    for (auto& lp : net->L) if (lp.s2d) {
      if (!f) f = &lp;
      else if (f->q.H != lp.q.H || f->q.W != lp.q.W || f->g.padT != lp.g.padT || f->g.padL != lp.g.padL) agree = false;
    }
    if (!agree) { for (auto& lp : net->L) if (lp.s2d) lp.tc = lp.s2d = false; f = nullptr; }
    if (f) { net->H4 = f->q.H; net->W4 = f->q.W; }
    net->obs_feats = net->H4 * net->W4 * 64;
  }
  // a tensor-core layer reads its source in batch-planar form: features must come in chunks of 8 (always true for
  // the shapes accepted above; a dense layer after an uncovered odd-width layer falls back)
  for (auto& lp : net->L) if (lp.tc && lp.d.src != 0 && net->tsize[lp.d.src] % 8) lp.tc = false;
  // workspace
  size_t w = 0;
  const int nt = desc->n_layers + 1;
  net->out_off.assign(nt, 0); net->gout_off.assign(nt, 0);
  net->obp_off.assign(nt, 0); net->gbp_off.assign(nt, 0);
  net->plane_elems.assign(nt, 0);
  for (int t = 1; t < nt; t++) {
    size_t bytes = align_up((size_t)max_batch * net->tsize[t] * sizeof(float), 256);
    net->out_off[t] = w; w += bytes;
    net->gout_off[t] = w; w += bytes;
    if (net->tsize[t] % 8 == 0) {
      long long pe = (long long)net->tsize[t] * net->pitch;
      net->plane_elems[t] = pe;
      size_t pbytes = align_up((size_t)pe * 2 * sizeof(uint16_t), 256);   // hi + lo
      net->obp_off[t] = w; w += pbytes;
      net->gbp_off[t] = w; w += pbytes;
    }
  }
  net->obs_bp_off = w; w += align_up((size_t)net->obs_feats * net->pitch * sizeof(uint16_t) + 256, 256);
  long long be = 0;
  for (auto& lp : net->L) if (lp.tc) {
    net->any_tc = true;
    lp.blob_off = be; be += (long long)lp.K * lp.N;
    net->blob_segs.push_back(bp::BlobSeg{lp.w_off, lp.blob_off, lp.K, lp.N, lp.s2d ? lp.k4 : 0});
  }
  net->blob_elems = (long long)align_up((size_t)be, 128);
  net->blob_off = w; w += align_up((size_t)net->blob_elems * 2 * sizeof(uint16_t), 256);
  for (auto& lp : net->L) if (lp.tc) {
    if (lp.d.kind == XTB_CONV) {
      lp.part_off = w; w += align_up((size_t)kSMs * lp.R * 128 * lp.N * sizeof(float), 256);
    } else {
    }
    lp.dbpart_off = w; w += align_up((size_t)kSMs * 64 * sizeof(float), 256);
  }
  for (auto& lp : net->L) if (lp.tc && lp.d.kind == XTB_CONV) {
    build_conv_tables(lp);
    auto place = [&](size_t bytes) { size_t o = w; w += align_up(bytes, 256); return o; };
    lp.fwd_st_off = place(lp.fwd_st.size() * sizeof(bp::StageEnt)); lp.fwd_un_off = place(lp.fwd_un.size() * sizeof(bp::UnitEnt));
    lp.dg_st_off = place(lp.dg_st.size() * sizeof(bp::StageEnt)); lp.dg_un_off = place(lp.dg_un.size() * sizeof(bp::UnitEnt));
    lp.wg_off = place(lp.wg_tab.size() * sizeof(bp::WgEnt));
  }
  // split-K partial sums of a dense forward: n_z slabs of [round16(B)][N] with n_z <= 148 / (N tiles * batch tiles) + 1
  net->splitk_off = w;
  {
    size_t need = 0;
    for (auto& lp : net->L) if (lp.tc && lp.d.kind == XTB_DENSE) {
      size_t rows = (size_t)(kSMs / (lp.N / lp.n_fwd) + 1) * 128 + net->pitch;
      need = std::max(need, rows * lp.N * sizeof(float));
    }
    w += align_up(need + 256, 256);
  }
  net->zeros_off = w; w += 4096;
  // per-block parameter-gradient slabs of the fused PPO heads kernel (K <= 512 hidden units, A <= 8 actions)
  net->heads_part_off = w; w += align_up((size_t)kSMs * (512 * 8 + 3 * 512 + 16) * sizeof(float), 256);
  net->segs_off = w; w += align_up(sizeof(bp::BlobSeg) * XTB_MAX_LAYERS, 256);
  net->ws_bytes = w;
  net->f32_ok.assign(nt, 0); net->bp_ok.assign(nt, 0); net->gf32_ok.assign(nt, 0); net->gbp_ok.assign(nt, 0);
  *out = net;
  return XTB_OK;
}

static void drop_graphs_of(const void* obj);
static void drop_step_graphs_of(const void* obj);
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

static cudaError_t ensure_kernel_attrs();
extern "C" int xtb_net_sync_weights(xtb_net* net, void* stream);
extern "C" int xtb_net_bind_stream(xtb_net* net, float* params, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
  if (!net || !params || !workspace) return fail(XTB_ERR_ARG, "xtb_net_bind: null pointer");
  if (workspace_bytes < net->ws_bytes) return fail(XTB_ERR_ARG, "workspace too small: %zu < %zu", workspace_bytes, net->ws_bytes);
  net->params = params; net->grads = grads; net->ws = (char*)workspace;
  cudaStream_t st = S(stream);
  { cudaError_t ea = ensure_kernel_attrs(); if (ea != cudaSuccess) return fail(XTB_ERR_CUDA, "kernel attributes: %s", cudaGetErrorString(ea)); }
  // Planes start as zeros: rows beyond the current batch are read (never used) by full-tile operand copies and
  // must be finite; the zero buffer feeds out-of-image chunks of padded weight-gradient operands.
  CUDA_TRY(cudaMemsetAsync(net->ws, 0, net->ws_bytes, st));
  if (!net->blob_segs.empty())
    CUDA_TRY(cudaMemcpyAsync(net->ws + net->segs_off, net->blob_segs.data(), net->blob_segs.size() * sizeof(bp::BlobSeg),
                             cudaMemcpyHostToDevice, st));
  for (const auto& lp : net->L) if (lp.tc && lp.d.kind == XTB_CONV) {
    auto up = [&](size_t off, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(net->ws + off, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess; };
    CUDA_TRY(up(lp.fwd_st_off, lp.fwd_st.data(), lp.fwd_st.size() * sizeof(bp::StageEnt)));
    CUDA_TRY(up(lp.fwd_un_off, lp.fwd_un.data(), lp.fwd_un.size() * sizeof(bp::UnitEnt)));
    CUDA_TRY(up(lp.dg_st_off, lp.dg_st.data(), lp.dg_st.size() * sizeof(bp::StageEnt)));
    CUDA_TRY(up(lp.dg_un_off, lp.dg_un.data(), lp.dg_un.size() * sizeof(bp::UnitEnt)));
    CUDA_TRY(up(lp.wg_off, lp.wg_tab.data(), lp.wg_tab.size() * sizeof(bp::WgEnt)));
  }
  std::fill(net->f32_ok.begin(), net->f32_ok.end(), 0); std::fill(net->bp_ok.begin(), net->bp_ok.end(), 0);
  std::fill(net->gf32_ok.begin(), net->gf32_ok.end(), 0); std::fill(net->gbp_ok.begin(), net->gbp_ok.end(), 0);
  net->obs_bp_ok = false;
  int rc = xtb_net_sync_weights(net, stream);
  if (rc) return rc;
  // the segment table came from pageable host memory of this call: do not return before it is on the device
  CUDA_TRY(cudaStreamSynchronize(st));
  return XTB_OK;
}
extern "C" int xtb_net_bind(xtb_net* net, float* params, float* grads, void* workspace, size_t workspace_bytes) {
  return xtb_net_bind_stream(net, params, grads, workspace, workspace_bytes, nullptr);
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
// fp32 GEMM dispatch
// ------------------------------------------------------------------------------------------
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
// tensor-core launches
// ------------------------------------------------------------------------------------------
static int tmem_cols_for(int cols) { int c = 32; while (c < cols) c <<= 1; return c; }

// opt-in to the large dynamic shared-memory carve-out, once per process and outside any stream capture
static cudaError_t ensure_kernel_attrs() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(bp::bp_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(bp::bp_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(bp::bp_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(bp::bp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(xtb::ppo_heads_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(xtb::ppo_heads_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(xtb::ppo_heads_kernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(xtb::ppo_heads_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem)) != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

template <int KIND>
static cudaError_t launch_rows(bp::RowsArgs& a, cudaStream_t st) {
  auto kern = bp::bp_rows_kernel<KIND>;
  { cudaError_t e0 = ensure_kernel_attrs(); if (e0 != cudaSuccess) return e0; }
  const int wres_bytes = a.w_res ? (int)align_up((size_t)2 * a.w_res_chunks * a.w_pitch * 16, 128) : 0;
  const int stage_bytes = bp::RW_STAGE_A + (a.w_res ? 0 : bp::RW_STAGE_B);
  const int tab_bytes = a.mode == 2 ? 0 : (int)align_up((size_t)a.n_stage_ents * sizeof(bp::StageEnt) + (size_t)a.n_units * sizeof(bp::UnitEnt), 128);
  int n_stages = std::min(bp::RW_MAX_STAGES, (kMaxDynSmem - 128 - wres_bytes - tab_bytes) / stage_bytes);
  if (n_stages < 2) return cudaErrorInvalidConfiguration;
  const int smem = 128 + wres_bytes + n_stages * stage_bytes + tab_bytes;
  const int acc = (KIND == 2 ? a.N : 2 * a.N);
  const int total = a.n_units * a.n_btiles;
  const int grid = std::min(total, kSMs);
  XLAUNCH(kern, grid, bp::RW_THREADS, smem, st, a, n_stages, stage_bytes, wres_bytes, tmem_cols_for(bp::RW_M_WARPS * acc));
  return cudaPeekAtLastError();
}

static cudaError_t launch_wgrad(const bp::WgradArgs& a, int grid, cudaStream_t st) {
  { cudaError_t e0 = ensure_kernel_attrs(); if (e0 != cudaSuccess) return e0; }
  const int smem = 128 + bp::WG_STAGES * bp::WG_STAGE + (a.mode == 0 ? (int)align_up((size_t)a.n_opix * a.R * sizeof(bp::WgEnt), 128) : 0);
  const int cols = (a.mode == 0 ? a.R : 1) * a.N;
  XLAUNCH(bp::bp_wgrad_kernel, grid, bp::WG_THREADS, smem, st, a, tmem_cols_for(cols));
  return cudaPeekAtLastError();
}

// forward of a tensor-core layer.  want_f32: also store the fp32 row-major copy; want_bp: store the planes
static cudaError_t tc_forward(xtb_net* net, int i, int B, bool want_f32, bool want_bp, cudaStream_t st, long long* launches) {
  const LayerPlan& lp = net->L[i];
  const int t = i + 1;
  bp::RowsArgs a;
  memset(&a, 0, sizeof a);
  a.a = lp.d.src == 0 ? obs_bp(net) : out_bp(net, lp.d.src);
  a.a_split = lp.d.src != 0;
  a.w_hi = blob_hi(net, lp); a.w_lo = a.w_hi + net->blob_elems; a.w_pitch = lp.K;
  a.B = B; a.n_btiles = (B + 127) / 128; a.N = lp.n_fwd;
  a.out = want_bp ? out_bp(net, t) : no_bp();
  a.out_f32 = want_f32 ? (float*)(net->ws + net->out_off[t]) : nullptr;
  a.ld_f32 = net->tsize[t];
  a.bias = net->params + lp.b_off;
  a.alpha = lp.d.src == 0 ? net->desc.scale : 1.f;
  a.act = lp.d.act;
  *launches = 1;
  if (lp.d.kind == XTB_CONV) {
    a.mode = 0;
    a.stages = (const bp::StageEnt*)(net->ws + lp.fwd_st_off); a.units = (const bp::UnitEnt*)(net->ws + lp.fwd_un_off);
    a.n_stage_ents = (int)lp.fwd_st.size();
    a.w_res = lp.w_res; a.w_res_chunks = lp.N / 8;
    a.n_units = lp.q.OH * lp.q.OW;
    return launch_rows<0>(a, st);
  }
  a.mode = 2;
  a.kchunks = lp.K / 8; a.n_ntiles = lp.N / lp.n_fwd;
  const int tiles = a.n_ntiles * a.n_btiles;
  int nz = 1;
  a.kc_split = a.kchunks;
  if (tiles < kSMs && a.kchunks >= 32) {
    // as many K slices as fit in ONE wave of CTAs (rounding up instead left a few CTAs with two slices: twice the latency)
    int want = std::min(kSMs / tiles, a.kchunks / 8);
    if (want > 1) {
      a.kc_split = ((a.kchunks + want - 1) / want + 7) / 8 * 8;     // whole 64-element stages per split
      nz = (a.kchunks + a.kc_split - 1) / a.kc_split;
    }
  }
  a.n_units = a.n_ntiles * nz;
  if (nz == 1) return launch_rows<0>(a, st);
  a.part = (float*)(net->ws + net->splitk_off);
  const int b_pad = (B + 15) & ~15;
  a.part_z = (long long)b_pad * lp.N; a.ld_part = lp.N;
  cudaError_t e = launch_rows<1>(a, st);
  if (e != cudaSuccess) return e;
  long long pieces = (long long)(lp.N / 8) * b_pad * bp::FIN_ZL;
  XLAUNCH(bp::bp_splitk_finish_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, (const float*)a.part, nz, a.part_z, B, lp.N,
          a.bias, lp.d.act, a.out_f32, a.out);
  *launches = 2;
  return cudaPeekAtLastError();
}

// data gradient of tensor-core layer i into the planes of its source tensor
static cudaError_t tc_dgrad(xtb_net* net, int i, int B, int accumulate, float* db_part, cudaStream_t st) {
  const LayerPlan& lp = net->L[i];
  const int t = i + 1, s = lp.d.src;
  bp::RowsArgs a;
  memset(&a, 0, sizeof a);
  a.a = gout_bp(net, t); a.a_split = 1;
  a.w_hi = blob_hi(net, lp); a.w_lo = a.w_hi + net->blob_elems; a.w_pitch = lp.K;
  a.B = B; a.n_btiles = (B + 127) / 128; a.N = lp.n_dg;
  a.out = gout_bp(net, s); a.src = out_bp(net, s); a.src_act = lp.src_act; a.accumulate = accumulate; a.db_part = db_part;
  if (lp.d.kind == XTB_CONV) {
    a.mode = 1;
    a.stages = (const bp::StageEnt*)(net->ws + lp.dg_st_off); a.units = (const bp::UnitEnt*)(net->ws + lp.dg_un_off);
    a.n_stage_ents = (int)lp.dg_st.size();
    a.w_res = lp.w_res; a.w_res_chunks = lp.N / 8;
    a.n_units = lp.q.H * lp.q.W;
  } else {
    a.mode = 2;
    a.kchunks = lp.N / 8; a.kc_split = a.kchunks; a.n_ntiles = lp.K / lp.n_dg;
    a.n_units = a.n_ntiles;
  }
  return launch_rows<2>(a, st);
}

// weight gradient of tensor-core layer i; conv layers queue their ordered reduction
static cudaError_t tc_wgrad(xtb_net* net, int i, int B, cudaStream_t st) {
  LayerPlan& lp = net->L[i];
  const int t = i + 1;
  bp::WgradArgs a;
  memset(&a, 0, sizeof a);
  a.x = lp.d.src == 0 ? obs_bp(net) : out_bp(net, lp.d.src);
  a.x_split = lp.d.src != 0;
  a.g = gout_bp(net, t);
  a.zeros = (const bp::bf16*)(net->ws + net->zeros_off);
  a.B = B; a.n_bsub = (B + bp::WG_KB - 1) / bp::WG_KB;
  if (lp.d.kind == XTB_CONV) {
    const ConvGeom& q = lp.q;
    a.mode = 0;
    a.tab = (const bp::WgEnt*)(net->ws + lp.wg_off);
    a.R = lp.R; a.N = lp.N; a.n_opix = q.OH * q.OW; a.n_ntiles = 1;
    a.part = (float*)(net->ws + lp.part_off);
    const int grid = std::min(kSMs, a.n_opix * a.n_bsub);
    bp::RedSeg r;
    memset(&r, 0, sizeof r);
    r.part = a.part; r.n_slabs = grid; r.slab = (long long)lp.R * 128 * lp.N; r.count = lp.R * 128 * lp.N; r.kind = 0;
    r.N = lp.N; r.C = q.C; r.KW = q.KW; r.mts = lp.mts; r.s2d_k4 = lp.s2d ? lp.k4 : 0;
    r.dst_off = lp.w_off; r.alpha = lp.d.src == 0 ? net->desc.scale : 1.f;
    net->pending.push_back(r);
    return launch_wgrad(a, grid, st);
  }
  a.mode = 1;
  a.N = std::min(lp.n_fwd, 64);
  a.x_chunks = lp.K / 8; a.r_tiles = (lp.K + 127) / 128; a.n_ntiles = lp.N / a.N; a.n_opix = 1; a.R = 1;
  a.dw = net->grads + lp.w_off; a.ldw = lp.N; a.k_rows = lp.K;
  return launch_wgrad(a, std::min(kSMs, a.r_tiles * a.n_ntiles), st);
}

// keep the weight blobs of the bound parameters current (after an optimiser step / set_weights)
extern "C" int xtb_net_sync_weights(xtb_net* net, void* stream) {
  if (!net || !net->ws || !net->params) return fail(XTB_ERR_STATE, "xtb_net_sync_weights: net not bound");
  if (net->blob_segs.empty()) return XTB_OK;
  long long mx = 0;
  for (const auto& s : net->blob_segs) mx = std::max(mx, (long long)(s.N / 8) * s.K);
  dim3 grid((unsigned)((mx + 127) / 128), (unsigned)net->blob_segs.size());
  XLAUNCH(bp::bp_wprep_kernel, grid, 128, 0, S(stream), (const float*)net->params, (const bp::BlobSeg*)(net->ws + net->segs_off),
          (bp::bf16*)(net->ws + net->blob_off), net->blob_elems);
  LAUNCH_CHECK();
  return XTB_OK;
}

#ifdef XTB_BP_TRACE
extern "C" int xtb_bp_trace_set(unsigned long long* buf) {
  CUDA_TRY(cudaMemcpyToSymbol(bp::g_bp_trace, &buf, sizeof buf));
  return XTB_OK;
}
#endif

// ---- representation changes ------------------------------------------------------------------
static int ensure_bp(xtb_net* net, int t, int B, bool grad, cudaStream_t st) {
  std::vector<char>& ok = grad ? net->gbp_ok : net->bp_ok;
  if (ok[t]) return XTB_OK;
  const std::vector<char>& f = grad ? net->gf32_ok : net->f32_ok;
  if (!f[t]) return fail(XTB_ERR_STATE, "tensor %d has no current %s", t, grad ? "gradient" : "value");
  const float* src = (const float*)(net->ws + (grad ? net->gout_off[t] : net->out_off[t]));
  const int F = net->tsize[t], b_pad = (B + 15) & ~15;
  long long pieces = (long long)(F / 8) * b_pad;
  XLAUNCH(bp::bp_split_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, src, B, F, grad ? gout_bp(net, t) : out_bp(net, t));
  LAUNCH_CHECK();
  ok[t] = 1;
  return XTB_OK;
}
static int ensure_f32(xtb_net* net, int t, int B, bool grad, cudaStream_t st) {
  std::vector<char>& ok = grad ? net->gf32_ok : net->f32_ok;
  if (ok[t]) return XTB_OK;
  const std::vector<char>& p = grad ? net->gbp_ok : net->bp_ok;
  if (!p[t]) return fail(XTB_ERR_STATE, "tensor %d has no current %s", t, grad ? "gradient" : "value");
  float* dst = (float*)(net->ws + (grad ? net->gout_off[t] : net->out_off[t]));
  const int F = net->tsize[t];
  long long pieces = (long long)(F / 8) * B;
  XLAUNCH(bp::bp_merge_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, grad ? gout_bp(net, t) : out_bp(net, t), B, F, dst);
  LAUNCH_CHECK();
  ok[t] = 1;
  return XTB_OK;
}

// Self-test of the three kernel forms on plain matrices (see tests/test_gpu_tc.py):
//   mode 0: C[M,N] = A[M,K]   * B[K,N]      forward   (rows kernel, weights MN-major; ksplit > 1: split-K partials)
//   mode 1: C[M,N] = A[M,K]   * Bt[N,K]^T   data grad (rows kernel, weights K-major)
//   mode 2: C[M,N] = At[K,M]^T * B[K,N]     weight grad (samples = K are the reduction axis)
extern "C" int xtb_tc_gemm_test(int mode, const float* a, const float* b, float* c, int M, int N, int K, int ksplit,
                                void* stream) {
  if (!a || !b || !c || M <= 0 || N <= 0 || K <= 0) return fail(XTB_ERR_ARG, "xtb_tc_gemm_test: bad argument");
  if (N % 16 || K % 16 || (mode == 2 && M % 16)) return fail(XTB_ERR_ARG, "xtb_tc_gemm_test: N, K (and M for mode 2) must be multiples of 16");
  cudaStream_t st = S(stream);
  // rows (samples): mode 0/1 -> M rows of A; mode 2 -> K rows of At and B
  const int rows = mode == 2 ? K : M;
  const int pitch = (rows + 15) / 16 * 16;
  const int fa = mode == 2 ? M : K;                 // features of the "activation" operand
  const int fb = N;                                  // mode 2: features of B
  const int nt = pick_tile(N);
  if (!nt) return fail(XTB_ERR_ARG, "N must be a multiple of 16");
  bp::bf16 *pa = nullptr, *pw = nullptr, *pc = nullptr; float* part = nullptr; float* bias = nullptr;
  const long long ea = (long long)fa * pitch;
  const long long ew = mode == 2 ? (long long)fb * pitch : (mode == 1 ? (long long)K * ((N + 15) / 16 * 16) : (long long)N * K);   // weight blob / second activation
  const long long ec = (long long)N * pitch;
  CUDA_TRY(cudaMalloc(&pa, 2 * ea * sizeof(bp::bf16)));
  CUDA_TRY(cudaMalloc(&pw, 2 * ew * sizeof(bp::bf16) + 256));
  CUDA_TRY(cudaMalloc(&pc, 2 * ec * sizeof(bp::bf16) + 4096));
  CUDA_TRY(cudaMalloc(&bias, (size_t)std::max(N, 64) * sizeof(float)));
  CUDA_TRY(cudaMemsetAsync(pa, 0, 2 * ea * sizeof(bp::bf16), st));
  CUDA_TRY(cudaMemsetAsync(pw, 0, 2 * ew * sizeof(bp::bf16) + 256, st));
  CUDA_TRY(cudaMemsetAsync(pc, 0, 2 * ec * sizeof(bp::bf16) + 4096, st));
  CUDA_TRY(cudaMemsetAsync(bias, 0, (size_t)std::max(N, 64) * sizeof(float), st));
  bp::BpT ta{pa, ea, pitch}, tw{pw, ew, mode == 2 ? pitch : K}, tcp{pc, ec, pitch};
  auto split = [&](const float* src, int B_, int F_, bp::BpT dst) {
    long long pieces = (long long)(F_ / 8) * ((B_ + 15) & ~15);
    XLAUNCH(bp::bp_split_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, src, B_, F_, dst);
  };
  cudaError_t e = cudaSuccess;
  int rc = XTB_OK;
  if (mode == 0 || mode == 1) {
    split(a, M, K, ta);
    // weight blob = batch-planar W^T with rows k: mode 0: b = W[K][N] -> rows K, features N.  mode 1: b = Bt[N][K]: the
    // kernel computes D[m, n] = sum_k A[m,k] Bt[n,k] with "W"[n][k] = Bt: rows = n (output), features = k (reduction)
    if (mode == 0) { tw.pitch = K; split(b, K, N, tw); }
    else { tw.pitch = (N + 15) / 16 * 16; split(b, N, K, tw); }
    bp::RowsArgs r;
    memset(&r, 0, sizeof r);
    r.a = ta; r.a_split = 1; r.w_hi = pw; r.w_lo = pw + ew; r.w_pitch = tw.pitch;
    r.mode = 2; r.B = M; r.n_btiles = (M + 127) / 128; r.N = nt; r.n_ntiles = N / nt;
    r.kchunks = K / 8; r.kc_split = r.kchunks;
    r.bias = bias; r.alpha = 1.f; r.act = 0;
    if (mode == 0) {
      int nz = 1;
      if (ksplit > 1) { r.kc_split = ((r.kchunks + ksplit - 1) / ksplit + 7) / 8 * 8; nz = (r.kchunks + r.kc_split - 1) / r.kc_split; }
      r.n_units = r.n_ntiles * nz;
      if (nz > 1) {
        CUDA_TRY(cudaMalloc(&part, (size_t)nz * M * N * sizeof(float)));
        r.part = part; r.part_z = (long long)M * N; r.ld_part = N;
        e = launch_rows<1>(r, st);
        long long pieces = (long long)(N / 8) * ((M + 15) & ~15) * bp::FIN_ZL;
        XLAUNCH(bp::bp_splitk_finish_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, (const float*)part, nz, r.part_z, M, N,
                (const float*)bias, 0, c, no_bp());
      } else {
        r.out_f32 = c; r.ld_f32 = N;
        e = launch_rows<0>(r, st);
      }
    } else {
      r.n_units = r.n_ntiles;
      r.out = tcp; r.src = tcp; r.src_act = 0;
      e = launch_rows<2>(r, st);
      long long pieces = (long long)(N / 8) * M;
      XLAUNCH(bp::bp_merge_kernel, (unsigned)((pieces + 127) / 128), 128, 0, st, tcp, M, N, c);
    }
  } else {
    split(a, K, M, ta);
    split(b, K, N, tw);
    bp::WgradArgs w;
    memset(&w, 0, sizeof w);
    w.x = ta; w.x_split = 1; w.g = tw; w.zeros = pc;
    w.mode = 1; w.N = nt; w.B = K; w.n_bsub = (K + bp::WG_KB - 1) / bp::WG_KB; w.n_opix = 1; w.R = 1;
    w.x_chunks = M / 8; w.r_tiles = (M + 127) / 128; w.n_ntiles = N / nt;
    w.dw = c; w.ldw = N; w.k_rows = M;
    e = launch_wgrad(w, std::min(kSMs, w.r_tiles * w.n_ntiles), st);
  }
  g_launches.fetch_add(4, std::memory_order_relaxed);
  cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(pa); cudaFree(pw); cudaFree(pc); cudaFree(bias); cudaFree(part);
  if (e != cudaSuccess) return fail(XTB_ERR_CUDA, "tc gemm launch: %s", cudaGetErrorString(e));
  if (e2 != cudaSuccess) return fail(XTB_ERR_CUDA, "tc gemm run: %s", cudaGetErrorString(e2));
  return rc;
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core layer ops
// ------------------------------------------------------------------------------------------
template <typename T>
static void conv_fwd(const LayerPlan& lp, const T* x, const int32_t* idx, const float* w, const float* b,
                     float alpha, float* out, int B, cudaStream_t st) {
  int M = B * lp.g.P;
  BRowMajor bl{w, lp.N};
  EpiBiasAct ep{out, b, alpha, lp.d.act, lp.N, nullptr, 0};
  if (lp.pad) { AIm2col<T, true> al{x, idx, lp.g, lp.koff, lp.kyx}; launch_gemm(al, bl, ep, M, lp.N, lp.K, false, st); }
  else { AIm2col<T, false> al{x, idx, lp.g, lp.koff, lp.kyx}; launch_gemm(al, bl, ep, M, lp.N, lp.K, false, st); }
}
template <typename T>
static void dense_fwd(const LayerPlan& lp, const T* x, const int32_t* idx, const float* w, const float* b,
                      float alpha, float* out, int B, cudaStream_t st) {
  ADense<T> al{x, idx, lp.K};
  BRowMajor bl{w, lp.N};
  EpiBiasAct ep{out, b, alpha, lp.d.act, lp.N, nullptr, 0};
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
// uint8 frame decode (+ minibatch gather) into the space-to-depth observation canvas
static int op_decode(xtb_net* net, const void* obs, const int32_t* idx, int B, cudaStream_t st) {
  if (net->obs_bp_ok) return XTB_OK;
  const LayerPlan* first = nullptr;
  for (const auto& lp : net->L) if (lp.s2d && use_tc(lp)) { first = &lp; break; }
  if (!first) return XTB_OK;
  dim3 grid(4 * net->H4, (B + bp::DEC_SAMPLES - 1) / bp::DEC_SAMPLES);
  const int smem = bp::DEC_SAMPLES * (net->desc.in_w * 4 + 8);
  XLAUNCH(bp::bp_decode_s2d_kernel, grid, bp::DEC_THREADS, smem, st, (const uint8_t*)obs, idx, B, net->desc.in_h, net->desc.in_w, net->H4,
          net->W4, first->g.padT, first->g.padL, obs_bp(net));
  LAUNCH_CHECK();
  net->obs_bp_ok = true;
  return XTB_OK;
}

// forward of layer i; tc_allowed = parameters are the bound ones (their blobs are current)
static int op_forward(xtb_net* net, int i, const float* P, bool tc_allowed, const void* obs, const int32_t* idx, int B,
                      bool want_f32, cudaStream_t st) {
  const LayerPlan& lp = net->L[i];
  const int t = i + 1;
  float* out = (float*)(net->ws + net->out_off[t]);
  const float* w = P + lp.w_off;
  const float* b = P + lp.b_off;
  if (tc_allowed && use_tc(lp)) {
    int rc = lp.d.src == 0 ? op_decode(net, obs, idx, B, st) : ensure_bp(net, lp.d.src, B, false, st);
    if (rc) return rc;
    long long nl = 0;
    cudaError_t te = tc_forward(net, i, B, want_f32, true, st, &nl);
    if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 forward launch (layer %d): %s", i, cudaGetErrorString(te));
    g_launches.fetch_add(nl, std::memory_order_relaxed);
    net->bp_ok[t] = 1; net->f32_ok[t] = want_f32 ? 1 : 0;
    return XTB_OK;
  }
  if (lp.d.src == 0) {
    float alpha = net->desc.scale;
    if (net->desc.input_u8) {
      if (lp.d.kind == XTB_CONV) conv_fwd<uint8_t>(lp, (const uint8_t*)obs, idx, w, b, alpha, out, B, st);
      else dense_fwd<uint8_t>(lp, (const uint8_t*)obs, idx, w, b, alpha, out, B, st);
    } else {
      if (lp.d.kind == XTB_CONV) conv_fwd<float>(lp, (const float*)obs, idx, w, b, alpha, out, B, st);
      else dense_fwd<float>(lp, (const float*)obs, idx, w, b, alpha, out, B, st);
    }
  } else {
    int rc = ensure_f32(net, lp.d.src, B, false, st);
    if (rc) return rc;
    const float* x = (const float*)(net->ws + net->out_off[lp.d.src]);
    if (lp.d.kind == XTB_CONV) conv_fwd<float>(lp, x, nullptr, w, b, 1.f, out, B, st);
    else dense_fwd<float>(lp, x, nullptr, w, b, 1.f, out, B, st);
  }
  LAUNCH_CHECK();
  net->f32_ok[t] = 1; net->bp_ok[t] = 0;
  return XTB_OK;
}

static int op_wgrad(xtb_net* net, int i, const void* obs, const int32_t* idx, int B, cudaStream_t st, bool bias_done = false) {
  const LayerPlan& lp = net->L[i];
  int t = i + 1;
  float* dw = net->grads + lp.w_off;
  float* db = net->grads + lp.b_off;
  if (use_tc(lp)) {
    int rc = ensure_bp(net, t, B, true, st);
    if (rc) return rc;
    rc = lp.d.src == 0 ? op_decode(net, obs, idx, B, st) : ensure_bp(net, lp.d.src, B, false, st);
    if (rc) return rc;
    cudaError_t te = tc_wgrad(net, i, B, st);
    if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 wgrad launch (layer %d): %s", i, cudaGetErrorString(te));
    LAUNCH_CHECK();
    if (!bias_done) {   // bias gradient = column sums of dY over samples (and pixels); from the fp32 copy when it is current
      if (net->gf32_ok[t]) {
        int Mb = lp.d.kind == XTB_CONV ? B * lp.g.P : B;
        dim3 gridb((lp.N + 31) / 32, (Mb + 1023) / 1024);
        XLAUNCH(colsum_kernel, gridb, 256, 0, st, (const float*)(net->ws + net->gout_off[t]), Mb, lp.N, db);
      } else {
        XLAUNCH(bp::bp_colsum_kernel, (net->tsize[t] / 8 + 3) / 4, 128, 0, st, gout_bp(net, t), B, net->tsize[t], lp.N, db);
      }
      LAUNCH_CHECK();
    }
    return XTB_OK;
  }
  int rc = ensure_f32(net, t, B, true, st);
  if (rc) return rc;
  const float* dy = (const float*)(net->ws + net->gout_off[t]);
  bool need_colsum = false;
  if (lp.d.src == 0) {
    float alpha = net->desc.scale;
    if (net->desc.input_u8) {
      if (lp.d.kind == XTB_CONV) conv_wgrad<uint8_t>(lp, (const uint8_t*)obs, idx, dy, alpha, dw, B, st, !bias_done);
      else dense_wgrad<uint8_t>(lp, (const uint8_t*)obs, idx, dy, alpha, dw, B, st, !bias_done);
    } else {
      if (lp.d.kind == XTB_CONV) conv_wgrad<float>(lp, (const float*)obs, idx, dy, alpha, dw, B, st, !bias_done);
      else dense_wgrad<float>(lp, (const float*)obs, idx, dy, alpha, dw, B, st, !bias_done);
    }
    need_colsum = alpha != 1.f;
  } else {
    rc = ensure_f32(net, lp.d.src, B, false, st);
    if (rc) return rc;
    const float* x = (const float*)(net->ws + net->out_off[lp.d.src]);
    if (lp.d.kind == XTB_CONV) conv_wgrad<float>(lp, x, nullptr, dy, 1.f, dw, B, st, !bias_done);
    else dense_wgrad<float>(lp, x, nullptr, dy, 1.f, dw, B, st, !bias_done);
  }
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
// fuse_db: the epilogue also produces the bias gradient of the layer behind the source tensor (ordered partial sums)
static int op_dgrad(xtb_net* net, int i, int acc, int B, cudaStream_t st, bool fuse_db = false) {
  const LayerPlan& lp = net->L[i];
  int t = i + 1, s = lp.d.src;
  if (use_tc(lp)) {
    int rc = ensure_bp(net, t, B, true, st);
    if (rc) return rc;
    if (lp.src_act != 0) { rc = ensure_bp(net, s, B, false, st); if (rc) return rc; }
    if (acc) { rc = ensure_bp(net, s, B, true, st); if (rc) return rc; }
    float* dbp = fuse_db ? (float*)(net->ws + lp.dbpart_off) : nullptr;
    cudaError_t te = tc_dgrad(net, i, B, acc, dbp, st);
    if (te != cudaSuccess) return fail(XTB_ERR_CUDA, "tcgen05 dgrad launch (layer %d): %s", i, cudaGetErrorString(te));
    LAUNCH_CHECK();
    if (fuse_db) {
      const int units = lp.d.kind == XTB_CONV ? lp.q.H * lp.q.W : lp.K / lp.n_dg;
      bp::RedSeg r;
      memset(&r, 0, sizeof r);
      r.part = dbp; r.n_slabs = std::min(kSMs, units * ((B + 127) / 128)); r.slab = lp.n_dg; r.count = lp.n_dg; r.kind = 1;
      r.dst_off = net->L[s - 1].b_off; r.alpha = 1.f;
      net->pending.push_back(r);
    }
    net->gbp_ok[s] = 1; net->gf32_ok[s] = 0;
    return XTB_OK;
  }
  int rc = ensure_f32(net, t, B, true, st);
  if (rc) return rc;
  rc = ensure_f32(net, s, B, false, st);
  if (rc) return rc;
  if (acc) { rc = ensure_f32(net, s, B, true, st); if (rc) return rc; }
  const float* dy = (const float*)(net->ws + net->gout_off[t]);
  const float* x = (const float*)(net->ws + net->out_off[s]);
  float* gsrc = (float*)(net->ws + net->gout_off[s]);
  const float* w = net->params + lp.w_off;
  if (lp.d.kind == XTB_CONV) {
    ADgrad al{dy, lp.g, lp.dkyx, lp.dco, lp.sshift};
    BConvDgrad bl{w, lp.wk, lp.N};
    EpiDgrad ep{gsrc, x, lp.src_act, lp.g.C, acc, nullptr, 0};
    launch_gemm(al, bl, ep, B * lp.g.H * lp.g.W, lp.g.C, lp.Kd, false, st);
  } else {
    ADense<float> al{dy, nullptr, lp.N};
    BTransposed bl{w, lp.N};
    EpiDgrad ep{gsrc, x, lp.src_act, lp.K, acc, nullptr, 0};
    launch_gemm(al, bl, ep, B, lp.K, lp.N, false, st);
  }
  LAUNCH_CHECK();
  net->gf32_ok[s] = 1; net->gbp_ok[s] = 0;
  return XTB_OK;
}

// queued ordered reductions (conv weight gradients, fused bias gradients) -> flat gradient bucket, one launch
static int flush_reductions(xtb_net* net, cudaStream_t st) {
  if (net->pending.empty()) return XTB_OK;
  bp::RedSegs segs;
  memset(&segs, 0, sizeof segs);
  int mx = 0;
  if (net->pending.size() > bp::RED_MAX) return fail(XTB_ERR_STATE, "too many pending reductions");
  for (size_t k = 0; k < net->pending.size(); k++) { segs.s[k] = net->pending[k]; mx = std::max(mx, net->pending[k].count); }
  dim3 grid((mx + 127) / 128, (unsigned)net->pending.size());
  XLAUNCH(bp::grad_reduce_kernel, grid, 256, 0, st, segs, net->grads);
  LAUNCH_CHECK();
  net->pending.clear();
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// forward / backward
// ------------------------------------------------------------------------------------------
// want_f32_mask: bit t set = tensor t is needed in fp32 row-major form (all tensors for the public entry point)
static int net_forward_impl(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                            int batch, void* stream, unsigned skip_mask, unsigned want_f32_mask);
extern "C" int xtb_net_forward(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                               int batch, void* stream) {
  return net_forward_impl(net, params, obs, gather_idx, batch, stream, 0u, ~0u);
}
static int net_forward_impl(xtb_net* net, const float* params, const void* obs, const int32_t* gather_idx,
                            int batch, void* stream, unsigned skip_mask, unsigned want_f32_mask) {
  if (!net || !net->ws) return fail(XTB_ERR_STATE, "xtb_net_forward: net not bound");
  if (batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "batch %d out of range (max %d)", batch, net->max_batch);
  if (!obs) return fail(XTB_ERR_ARG, "obs is null");
  const float* P = params ? params : net->params;
  const bool tc_allowed = (P == net->params);   // foreign parameters have no weight blobs: fp32 kernels
  cudaStream_t st = S(stream);
  const int nl = (int)net->L.size();
  net->obs_bp_ok = false;
  for (int t = 1; t <= nl; t++) { net->f32_ok[t] = net->bp_ok[t] = 0; }
  for (int i = 0; i < nl; i++) {
    if (skip_mask & (1u << i)) continue;
    const LayerPlan& lp = net->L[i];
    // dense layers write fp32 from the epilogue when asked; conv maps are merged afterwards (the only fp32 readers of
    // conv maps are uncovered layers and the public API)
    bool direct = lp.d.kind == XTB_DENSE && ((want_f32_mask >> (i + 1)) & 1u);
    // ... or when a layer outside the tensor-core path consumes it (exact fp32 instead of hi + lo)
    if (lp.d.kind == XTB_DENSE)
      for (int j = i + 1; j < nl; j++)
        if (net->L[j].d.src == i + 1 && !(skip_mask & (1u << j)) && !(tc_allowed && use_tc(net->L[j]))) direct = true;
    int rc = op_forward(net, i, P, tc_allowed, obs, gather_idx, batch, direct, st);
    if (rc) return rc;
  }
  for (int t = 1; t <= nl; t++) {
    if (!((want_f32_mask >> t) & 1u) || (skip_mask & (1u << (t - 1)))) continue;
    int rc = ensure_f32(net, t, batch, false, st);
    if (rc) return rc;
  }
  return XTB_OK;
}

static int net_backward_impl(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                             const int32_t* head_tensors, int n_heads, void* stream, unsigned skip_mask, bool zero_grads,
                             unsigned bias_done_tensors = 0u, unsigned heads_bp_mask = 0u, xtb_comm* comm = nullptr);
extern "C" int xtb_net_backward(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                                const int32_t* head_tensors, int n_heads, void* stream) {
  return net_backward_impl(net, obs, gather_idx, batch, head_tensors, n_heads, stream, 0u, true);
}
// head_tensors: tensors whose gradient was filled by the caller: fp32 row-major, or (bit set in heads_bp_mask) planes
static int net_backward_impl(xtb_net* net, const void* obs, const int32_t* gather_idx, int batch,
                             const int32_t* head_tensors, int n_heads, void* stream, unsigned skip_mask, bool zero_grads,
                             unsigned bias_done_tensors, unsigned heads_bp_mask, xtb_comm* comm) {
  if (!net || !net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_net_backward: net not bound (grads required)");
  if (batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "batch out of range");
  cudaStream_t st = S(stream);
  const int nl = (int)net->L.size();
  std::vector<char> has_grad(nl + 1, 0), written(nl + 1, 0);
  long long early_off = 0, early_cnt = 0;
  for (int t = 1; t <= nl; t++) net->gf32_ok[t] = net->gbp_ok[t] = 0;
  if (zero_grads) net->pending.clear();      // a fused loss kernel (zero_grads == false) has queued its own reductions
  for (int h = 0; h < n_heads; h++) {
    int t = head_tensors[h];
    if (t < 1 || t > nl) return fail(XTB_ERR_ARG, "bad head tensor %d", t);
    has_grad[t] = 1; written[t] = 1;
    if ((heads_bp_mask >> t) & 1u) net->gbp_ok[t] = 1; else net->gf32_ok[t] = 1;
  }
  if (zero_grads) CUDA_TRY(cudaMemsetAsync(net->grads, 0, net->n_params * sizeof(float), st));
  // bias-gradient fusion: the bias gradient of the layer producing tensor s is the column sum of gout(s); when s has
  // exactly one consumer and that consumer's tensor-core data-gradient tile spans exactly the bias vector, its
  // epilogue accumulates the column sums (ordered partial sums, no atomics) and the separate pass is dropped
  std::vector<char> fuse_bias(nl + 1, 0);
  for (int s = 1; s <= nl; s++) {
    if (bias_done_tensors & (1u << s)) { fuse_bias[s] = 2; continue; }
    const LayerPlan& ps = net->L[s - 1];
    int consumers = 0, cj = -1;
    for (int j = 0; j < nl; j++) if (net->L[j].d.src == s && !(skip_mask & (1u << j))) { consumers++; cj = j; }
    if (consumers != 1) continue;
    const LayerPlan& c = net->L[cj];
    if (!use_tc(c)) continue;
    if (ps.d.kind == XTB_CONV ? c.n_dg == ps.N : (c.d.kind == XTB_DENSE && c.n_dg == ps.N && c.K == ps.N)) fuse_bias[s] = 1;
  }
  for (int i = nl - 1; i >= 0; i--) {
    if (skip_mask & (1u << i)) continue;
    const LayerPlan& lp = net->L[i];
    int t = i + 1;
    if (!has_grad[t]) continue;   // tensor does not influence the loss
    int rc = op_wgrad(net, i, obs, gather_idx, batch, st, fuse_bias[t] != 0);
    if (rc) return rc;
    // data parallel: a large dense weight gradient is final here (direct store) -- start summing it over ranks now, on
    // the communicator's side stream, while the rest of the backward pass runs
    if (comm && comm->world > 1 && early_cnt == 0 && lp.d.kind == XTB_DENSE && use_tc(lp) && (long long)lp.K * lp.N >= (1 << 16)) {
      early_off = lp.w_off; early_cnt = (long long)lp.K * lp.N;
      rc = comm_fork_allreduce(comm, net->grads + early_off, early_cnt, st);
      if (rc) return rc;
    }
    if (lp.d.src != 0) {
      int s = lp.d.src;
      rc = op_dgrad(net, i, written[s] ? 1 : 0, batch, st, fuse_bias[s] == 1);
      if (rc) return rc;
      written[s] = 1; has_grad[s] = 1;
    }
  }
  int rc = flush_reductions(net, st);
  if (rc || !comm || comm->world == 1) return rc;
  // the rest of the bucket: everything before and after the early range, as one NCCL group
  if (early_cnt == 0) return xtb_comm_allreduce(comm, net->grads, net->n_params, stream);
  int nrc = g_nccl.GroupStart();
  if (!nrc && early_off > 0) nrc = g_nccl.AllReduce(net->grads, net->grads, (size_t)early_off, kNcclFloat, kNcclSum, comm->comm, st);
  const long long tail = net->n_params - (early_off + early_cnt);
  if (!nrc && tail > 0)
    nrc = g_nccl.AllReduce(net->grads + early_off + early_cnt, net->grads + early_off + early_cnt, (size_t)tail, kNcclFloat, kNcclSum, comm->comm, st);
  int erc = g_nccl.GroupEnd();
  if (nrc || erc) return fail(XTB_ERR_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc ? nrc : erc));
  return comm_join(comm, st);
}

// Launch ONE kernel of one layer (0 = forward, 1 = weight gradient, 2 = data gradient, 3 = frame decode) on the
// tensors currently in the workspace: lets bench.py time the dominant kernel alone with CUDA events.
extern "C" int xtb_net_bench_layer(xtb_net* net, int layer, int which, const void* obs, const int32_t* gather_idx,
                                   int batch, void* stream) {
  if (!net || !net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_net_bench_layer: net not bound");
  if (layer < 0 || layer >= (int)net->L.size() || batch <= 0 || batch > net->max_batch) return fail(XTB_ERR_ARG, "bad layer/batch");
  cudaStream_t st = S(stream);
  // tensors the caller filled through xtb_net_tensor / xtb_net_tensor_grad are fp32 row-major
  for (int t = 1; t <= (int)net->L.size(); t++) {
    if (!net->f32_ok[t] && !net->bp_ok[t]) net->f32_ok[t] = 1;
    if (!net->gf32_ok[t] && !net->gbp_ok[t]) net->gf32_ok[t] = 1;
  }
  if (which == 0) return op_forward(net, layer, net->params, true, obs, gather_idx, batch, false, st);
  if (which == 1) { int rc = op_wgrad(net, layer, obs, gather_idx, batch, st, true); net->pending.clear(); return rc; }
  if (which == 2) {
    if (net->L[layer].d.src == 0) return fail(XTB_ERR_ARG, "layer reads the observation: no data gradient");
    int rc = op_dgrad(net, layer, 0, batch, st);
    net->pending.clear();
    return rc;
  }
  if (which == 3) { net->obs_bp_ok = false; return op_decode(net, obs, gather_idx, batch, st); }
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

extern "C" int xtb_dqn_td_loss_grad(const float* q, const float* q_next_target, const float* q_next_online, const int32_t* idx,
                                    const int32_t* action, const float* reward, const uint8_t* done, const float* disc, int batch,
                                    int adim, float gamma, float huber_delta, float inv_count, float* dq, float* y_out,
                                    float* loss_out, void* stream) {
  if (!q || !q_next_target || !action || !reward || !done || !dq || !loss_out) return fail(XTB_ERR_ARG, "xtb_dqn_td_loss_grad: null pointer");
  if (batch <= 0 || adim <= 0) return fail(XTB_ERR_ARG, "xtb_dqn_td_loss_grad: bad sizes");
  XLAUNCH(dqn_loss_kernel, (batch + 127) / 128, 128, 0, S(stream), q, q_next_target, q_next_online, idx, action, reward, done, disc,
          batch, adim, gamma, huber_delta, inv_count, dq, y_out, loss_out);
  LAUNCH_CHECK();
  return XTB_OK;
}
extern "C" int xtb_dqn_loss_grad(const float* q, const float* q_next_target, const float* q_next_online,
                                 const int32_t* action, const float* reward, const uint8_t* done, int batch,
                                 int adim, float gamma, float inv_count, float* dq, float* y_out,
                                 float* loss_out, void* stream) {
  return xtb_dqn_td_loss_grad(q, q_next_target, q_next_online, nullptr, action, reward, done, nullptr, batch, adim, gamma, 0.f,
                              inv_count, dq, y_out, loss_out, stream);
}
extern "C" int xtb_nstep_returns(const float* reward, const uint8_t* done, int n_env, int n_step, int n, float gamma, float* ret,
                                 float* disc, int32_t* last, uint8_t* done_n, void* stream) {
  if (!reward || !done || !ret || !disc || !last || !done_n) return fail(XTB_ERR_ARG, "xtb_nstep_returns: null pointer");
  if (n_env <= 0 || n_step <= 0 || n <= 0) return fail(XTB_ERR_ARG, "xtb_nstep_returns: bad sizes");
  XLAUNCH(nstep_kernel, (n_env * n_step + 127) / 128, 128, 0, S(stream), reward, done, n_env, n_step, n, gamma, ret, disc, last, done_n);
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
  float* mg = nullptr; float rms_rho = 0.f, rms_eps = 0.f;   // centred RMSProp instead of Adam when mg != NULL (m = ms)
  int* blk_seg = nullptr; long long* blk_beg = nullptr; int* blk_len = nullptr;
  double* norm_sq = nullptr; float* seg_scale = nullptr; AdamState* st = nullptr; AdamHyper* hyp = nullptr; unsigned int* ticket = nullptr;
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
  chk(cudaMalloc(&o->hyp, sizeof(AdamHyper)));
  chk(cudaMalloc(&o->ticket, sizeof(unsigned int)));
  if (e == cudaSuccess) {
    chk(cudaMemcpy(o->blk_seg, bseg.data(), o->n_blk * sizeof(int), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(o->blk_beg, bbeg.data(), o->n_blk * sizeof(long long), cudaMemcpyHostToDevice));
    chk(cudaMemcpy(o->blk_len, blen.data(), o->n_blk * sizeof(int), cudaMemcpyHostToDevice));
    chk(cudaMemset(o->norm_sq, 0, o->n_seg * sizeof(double)));
    chk(cudaMemcpy(o->st, &init, sizeof init, cudaMemcpyHostToDevice));
    chk(cudaMemset(o->ticket, 0, sizeof(unsigned int)));
    AdamHyper hy{lr, beta1, beta2, eps, clip};
    chk(cudaMemcpy(o->hyp, &hy, sizeof hy, cudaMemcpyHostToDevice));
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
  cudaFree(o->norm_sq); cudaFree(o->seg_scale); cudaFree(o->st); cudaFree(o->hyp); cudaFree(o->ticket);
  delete o;
}

static int adam_step_impl(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream, xtb_net* net);
extern "C" int xtb_adam_step(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream) {
  return adam_step_impl(o, params, grads, grad_scale, stream, nullptr);
}
// optimiser step on a network's bound parameters; the same kernel refreshes the weight blobs of its tensor-core layers
extern "C" int xtb_adam_step_net(xtb_adam* o, xtb_net* net, float grad_scale, void* stream) {
  if (!net || !net->ws || !net->params || !net->grads) return fail(XTB_ERR_STATE, "xtb_adam_step_net: net not bound");
  if (!o || o->count != net->n_params) return fail(XTB_ERR_ARG, "xtb_adam_step_net: optimiser/net size mismatch");
  return adam_step_impl(o, net->params, net->grads, grad_scale, stream, net);
}
static int adam_step_impl(xtb_adam* o, float* params, const float* grads, float grad_scale, void* stream, xtb_net* net) {
  if (!o || !params || !grads) return fail(XTB_ERR_ARG, "xtb_adam_step: null pointer");
  cudaStream_t st = S(stream);
  XLAUNCH(sqnorm_kernel, (o->n_blk + SQN_GROUP - 1) / SQN_GROUP, OPT_THREADS, 0, st, grads, o->blk_seg, o->blk_beg, o->blk_len, o->n_blk, o->norm_sq, o->ticket, o->st,
          (const AdamHyper*)o->hyp, o->seg_scale, o->n_seg, o->clip_mode, grad_scale);
  LAUNCH_CHECK();
  const bool blobs = net && !net->blob_segs.empty();
  if (o->mg) {
    XLAUNCH(rmsprop_kernel, o->n_blk, OPT_THREADS, 0, st, params, grads, o->m, o->mg, o->blk_seg, o->blk_beg, o->blk_len,
            o->seg_scale, (const AdamHyper*)o->hyp, o->rms_rho, o->rms_eps, blobs ? (const bp::BlobSeg*)(net->ws + net->segs_off) : nullptr,
            blobs ? (int)net->blob_segs.size() : 0, blobs ? (__nv_bfloat16*)(net->ws + net->blob_off) : nullptr, blobs ? net->blob_elems : 0LL);
    LAUNCH_CHECK();
    return XTB_OK;
  }
  XLAUNCH(adam_kernel, o->n_blk, OPT_THREADS, 0, st, params, grads, o->m, o->v, o->blk_seg, o->blk_beg, o->blk_len,
          o->seg_scale, o->st, (const AdamHyper*)o->hyp, blobs ? (const bp::BlobSeg*)(net->ws + net->segs_off) : nullptr,
          blobs ? (int)net->blob_segs.size() : 0, blobs ? (__nv_bfloat16*)(net->ws + net->blob_off) : nullptr, blobs ? net->blob_elems : 0LL);
  LAUNCH_CHECK();
  return XTB_OK;
}

extern "C" const float* xtb_adam_grad_norm(const xtb_adam* o) { return o ? &o->st->grad_norm : nullptr; }
extern "C" int xtb_opt_use_rmsprop(xtb_adam* o, float* mean_grad, float decay, float epsilon) {
  if (!o || !mean_grad) return fail(XTB_ERR_ARG, "xtb_opt_use_rmsprop: null pointer");
  if (!(decay > 0.f && decay < 1.f) || !(epsilon > 0.f)) return fail(XTB_ERR_ARG, "xtb_opt_use_rmsprop: decay in (0,1), epsilon > 0");
  drop_graphs_of(o);                       // captured steps baked the Adam kernel in
  {   // slot initial values of tf.train.RMSPropOptimizer: rms = ones, mg = zeros (one-time, synchronous)
    std::vector<float> ones((size_t)o->count, 1.f);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(o->m, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(mean_grad, 0, ones.size() * sizeof(float)));
  }
  o->mg = mean_grad; o->rms_rho = decay; o->rms_eps = epsilon;
  return XTB_OK;
}
extern "C" int xtb_adam_set_lr(xtb_adam* o, float lr) {
  if (!o) return fail(XTB_ERR_ARG, "null optimiser");
  o->lr = lr;
  // device-resident: captured graphs read it at replay time.  Ordered after everything already submitted.
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(&o->hyp->lr, &lr, sizeof lr, cudaMemcpyHostToDevice));
  return XTB_OK;
}

// ------------------------------------------------------------------------------------------
// fused PPO learner loop
// ------------------------------------------------------------------------------------------
// CUDA graphs cannot be captured on the legacy default stream, which is what a host that never creates streams
// (stream == NULL) runs on.  Such calls are moved onto a private non-blocking stream of the current device, fenced
// against the legacy stream with events on both sides, so the caller keeps default-stream ordering semantics.
struct EngineStream { cudaStream_t st = nullptr; cudaEvent_t in = nullptr, out = nullptr; };
static EngineStream g_engine_streams[64];
struct StreamScope {
  cudaStream_t st = nullptr;
  EngineStream* es = nullptr;
  int begin(void* stream, bool side_if_null) {
    st = S(stream);
    if (st || !side_if_null) return XTB_OK;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(XTB_ERR_ARG, "device index %d out of range", dev);
    EngineStream& e = g_engine_streams[dev];
    if (!e.st) {
      CUDA_TRY(cudaStreamCreateWithFlags(&e.st, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&e.in, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&e.out, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(e.in, nullptr));
    CUDA_TRY(cudaStreamWaitEvent(e.st, e.in, 0));
    es = &e; st = e.st;
    return XTB_OK;
  }
  int end() {
    if (!es) return XTB_OK;
    CUDA_TRY(cudaEventRecord(es->out, es->st));
    CUDA_TRY(cudaStreamWaitEvent(nullptr, es->out, 0));
    es = nullptr;
    return XTB_OK;
  }
};
static std::atomic<long long> g_graph_replays{0};
extern "C" long long xtb_graph_replay_count(void) { return g_graph_replays.load(); }

// A captured graph bakes every kernel argument: the key carries everything that can change them.
struct GraphKey {
  const void* net; const void* opt; const void* obs; const void* perm; const void* loss; int n, b, e;
  const void* ro[5]; float hp[4]; int pi_t, v_t, fuse, tc; const void* ws; const void* comm;
  bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
};
struct GraphVal { cudaGraphExec_t exec; long long kernels; };
static constexpr size_t kMaxCachedGraphs = 256;   // per cache; beyond it the cache is emptied (keys are buffer addresses)
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
      // fp32 row-major copies: the hidden tensors the fused heads read, or the head outputs the loss kernel reads
      unsigned want = fuse ? ((1u << lpi.d.src) | (1u << lv.d.src)) : ((1u << pi_t) | (1u << v_t));
      int rc = net_forward_impl(net, nullptr, ro->obs, idx, mb, stream, skip, want);
      if (rc) return rc;
      if (fuse) {
        CUDA_TRY(cudaMemsetAsync(net->grads, 0, net->n_params * sizeof(float), S(stream)));
        net->pending.clear();
        PpoHeadsArgs a;
        a.h_pi = (const float*)(net->ws + net->out_off[lpi.d.src]); a.h_v = (const float*)(net->ws + net->out_off[lv.d.src]);
        a.g_pi = (float*)(net->ws + net->gout_off[lpi.d.src]); a.g_v = (float*)(net->ws + net->gout_off[lv.d.src]);
        // hidden-layer gradients go straight into batch-planar planes when the hidden layer runs on tensor cores
        const bool bp_pi = use_tc(net->L[lpi.d.src - 1]) && net->plane_elems[lpi.d.src] > 0;
        const bool bp_v = use_tc(net->L[lv.d.src - 1]) && net->plane_elems[lv.d.src] > 0;
        a.gp_hi = bp_pi ? gout_bp(net, lpi.d.src).hi : nullptr; a.gp_lo = net->plane_elems[lpi.d.src];
        a.gv_hi = bp_v ? gout_bp(net, lv.d.src).hi : nullptr; a.gv_lo = net->plane_elems[lv.d.src];
        a.pitch = net->pitch;
        a.w_pi = net->params + lpi.w_off; a.b_pi = net->params + lpi.b_off; a.w_v = net->params + lv.w_off; a.b_v = net->params + lv.b_off;
        // the hidden layers' bias gradients (column sums of g) when they are dense and only feed the heads
        auto only_feeds_heads = [&](int tsr) { for (int j = 0; j < (int)net->L.size(); j++) if (net->L[j].d.src == tsr && !(skip & (1u << j))) return false; return true; };
        bool bh_pi_ok = net->L[lpi.d.src - 1].d.kind == XTB_DENSE && only_feeds_heads(lpi.d.src);
        bool bh_v_ok = net->L[lv.d.src - 1].d.kind == XTB_DENSE && only_feeds_heads(lv.d.src);
        unsigned bias_done = (bh_pi_ok ? (1u << lpi.d.src) : 0u) | ((lpi.d.src != lv.d.src && bh_v_ok) ? (1u << lv.d.src) : 0u);
        a.idx = idx; a.action = ro->action; a.old_logp = ro->old_logp; a.adv = ro->adv; a.old_v = ro->old_v; a.target_v = ro->target_v;
        a.logits_out = xtb_net_tensor(net, pi_t); a.v_out = xtb_net_tensor(net, v_t);
        a.B = mb; a.K = lpi.K; a.A = adim; a.act_pi = lpi.src_act; a.act_v = lv.src_act; a.shared = lpi.d.src == lv.d.src ? 1 : 0;
        a.hp = PpoHyperDev{hp->clip_ratio, hp->ent_coef, hp->vf_clip, hp->critic_coef}; a.inv_count = inv_world / mb;
        int blocks = std::max(1, std::min(148, (mb + 7) / 8));      // one sample per warp up to 1184 samples
        const int HK = lpi.K, nacc = HK * adim + 3 * HK + adim + 2;
        a.part = (float*)(net->ws + net->heads_part_off); a.slab = (nacc + 3) & ~3;
        size_t shb = (size_t)8 * nacc * sizeof(float);
        { cudaError_t ea = ensure_kernel_attrs(); if (ea != cudaSuccess) return fail(XTB_ERR_CUDA, "kernel attributes: %s", cudaGetErrorString(ea)); }
        if (kpl <= 2) XLAUNCH((ppo_heads_kernel<2, 8>), blocks, 256, shb, S(stream), a);
        else if (kpl <= 8 && adim <= 4) XLAUNCH((ppo_heads_kernel<8, 4>), blocks, 256, shb, S(stream), a);
        else if (kpl <= 8) XLAUNCH((ppo_heads_kernel<8, 8>), blocks, 256, shb, S(stream), a);
        else XLAUNCH((ppo_heads_kernel<16, 4>), blocks, 256, shb, S(stream), a);
        LAUNCH_CHECK();
        {   // ordered reduction of the per-block slabs (queued; runs with the other partial sums at the end of backward)
          auto seg = [&](int off, int count, long long dst_off, float* dst_ptr) {
            bp::RedSeg r;
            memset(&r, 0, sizeof r);
            r.part = a.part + off; r.n_slabs = blocks; r.slab = a.slab; r.count = count; r.kind = 1;
            r.dst_off = dst_off; r.alpha = 1.f; r.dst_ptr = dst_ptr;
            net->pending.push_back(r);
          };
          seg(0, HK * adim, lpi.w_off, nullptr);
          seg(HK * adim, HK, lv.w_off, nullptr);
          if (bh_pi_ok) seg(HK * adim + HK, HK, net->L[lpi.d.src - 1].b_off, nullptr);
          if (!a.shared && bh_v_ok) seg(HK * adim + 2 * HK, HK, net->L[lv.d.src - 1].b_off, nullptr);
          seg(HK * adim + 3 * HK, adim, lpi.b_off, nullptr);
          seg(HK * adim + 3 * HK + adim, 1, lv.b_off, nullptr);
          seg(HK * adim + 3 * HK + adim + 1, 1, 0, loss_per_step + step);
        }
        int srcs[2] = {lpi.d.src, lv.d.src};
        unsigned hbp = (bp_pi ? (1u << lpi.d.src) : 0u) | (bp_v ? (1u << lv.d.src) : 0u);   // the fused kernel wrote planes there
        rc = net_backward_impl(net, ro->obs, idx, mb, srcs, a.shared ? 1 : 2, stream, skip, false, bias_done, hbp, g_comm);
        if (rc) return rc;
      } else {
        rc = xtb_ppo_loss_grad(xtb_net_tensor(net, pi_t), xtb_net_tensor(net, v_t), idx, ro->action, ro->old_logp,
                               ro->adv, ro->old_v, ro->target_v, mb, adim, hp, inv_world / mb,
                               xtb_net_tensor_grad(net, pi_t), xtb_net_tensor_grad(net, v_t), loss_per_step + step, stream);
        if (rc) return rc;
        rc = net_backward_impl(net, ro->obs, idx, mb, heads, 2, stream, 0u, true, 0u, 0u, g_comm);
        if (rc) return rc;
      }
      if (g_grad_hook && !g_comm) {
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
  if (g_comm) inv_world = 1.f / g_comm->world;
  else if (g_grad_hook) {   // data-parallel: the hook sums gradients over ranks; every rank holds B/world samples
    int world = 1;
    int rc = g_grad_hook(g_grad_hook_user, nullptr, 0, nullptr);   // query: returns world size when grads == NULL
    if (rc > 0) world = rc;
    inv_world = 1.f / world;
  }
  if (!use_graph || (g_grad_hook && !g_comm))
    return ppo_train_launch(net, opt, ro, n_sample, batch_size, n_epoch, perm, hp, pi_tensor, v_tensor, loss_per_step, inv_world, stream);
  StreamScope sc;
  int src = sc.begin(stream, true);
  if (src) return src;
  GraphKey key;
  memset(&key, 0, sizeof key);
  key.net = net; key.opt = opt; key.obs = ro->obs; key.perm = perm; key.loss = loss_per_step;
  key.n = n_sample; key.b = batch_size; key.e = n_epoch;
  key.ro[0] = ro->action; key.ro[1] = ro->old_logp; key.ro[2] = ro->adv; key.ro[3] = ro->old_v; key.ro[4] = ro->target_v;
  key.hp[0] = hp->clip_ratio; key.hp[1] = hp->ent_coef; key.hp[2] = hp->vf_clip; key.hp[3] = hp->critic_coef;
  key.pi_t = pi_tensor; key.v_t = v_tensor; key.fuse = g_fuse_heads; key.tc = g_tc_mode; key.ws = net->ws; key.comm = g_comm;
  auto it = g_graphs.find(key);
  if (it == g_graphs.end()) {
    cudaStream_t st = sc.st;
    long long before = g_launches.load();
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = ppo_train_launch(net, opt, ro, n_sample, batch_size, n_epoch, perm, hp, pi_tensor, v_tensor, loss_per_step, inv_world, (void*)st);
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
    if (g_graphs.size() >= kMaxCachedGraphs) {
      for (auto& kv : g_graphs) cudaGraphExecDestroy(kv.second.exec);
      g_graphs.clear();
    }
    it = g_graphs.emplace(key, GraphVal{exec, captured}).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second.exec, sc.st));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);
  g_graph_replays.fetch_add(1, std::memory_order_relaxed);
  return sc.end();
}

// ------------------------------------------------------------------------------------------
// fused IMPALA / DQN learner steps (graph-captured like xtb_ppo_train; gradients all-reduced when a communicator is set)
// ------------------------------------------------------------------------------------------
struct StepKey {
  const void* p[12]; int i[8]; float f[4];
  bool operator<(const StepKey& o) const { return memcmp(this, &o, sizeof(StepKey)) < 0; }
};
static std::map<StepKey, GraphVal> g_step_graphs;

static void drop_step_graphs_of(const void* obj) {
  for (auto it = g_step_graphs.begin(); it != g_step_graphs.end();) {
    if (it->first.p[0] == obj || it->first.p[1] == obj || it->first.p[10] == obj || it->first.p[11] == obj) { cudaGraphExecDestroy(it->second.exec); it = g_step_graphs.erase(it); } else ++it;
  }
}

template <class F>
static int run_step_graph(const StepKey& key, int use_graph, void* stream, F&& launch) {
  if (!use_graph) return launch(stream);
  StreamScope sc;
  int src = sc.begin(stream, true);
  if (src) return src;
  auto it = g_step_graphs.find(key);
  if (it == g_step_graphs.end()) {
    cudaStream_t st = sc.st;
    long long before = g_launches.load();
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = launch((void*)st);
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
    if (g_step_graphs.size() >= kMaxCachedGraphs) {     // callers that pass fresh buffers every step must not leak executables
      for (auto& kv : g_step_graphs) cudaGraphExecDestroy(kv.second.exec);
      g_step_graphs.clear();
    }
    it = g_step_graphs.emplace(key, GraphVal{exec, captured}).first;
  }
  CUDA_TRY(cudaGraphLaunch(it->second.exec, sc.st));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);
  g_graph_replays.fetch_add(1, std::memory_order_relaxed);
  return sc.end();
}

// ImpalaCnnOpt.train (xt/model/impala/impala_cnn_opt.py:251-265): forward over n = k * step_len env-major samples,
// V-trace + summed losses (vtrace_kernel), backward, clip + Adam.  With a communicator the losses are sums over the
// GLOBAL batch, so no rescaling: gradients are summed over ranks.  loss_out: device float, accumulated (+=).
extern "C" int xtb_impala_train(xtb_net* net, xtb_adam* opt, const void* obs, const int32_t* gather_idx, const float* bp_logits,
                                const int32_t* action, const uint8_t* done, const float* reward, int n_sample, int step_len,
                                float gamma, int logit_tensor, int base_tensor, float* loss_out, int use_graph, void* stream) {
  if (!net || !opt || !obs || !bp_logits || !action || !done || !reward || !loss_out) return fail(XTB_ERR_ARG, "xtb_impala_train: null pointer");
  if (!net->ws || !net->grads) return fail(XTB_ERR_STATE, "xtb_impala_train: net not bound");
  const int nl = (int)net->L.size();
  if (logit_tensor < 1 || logit_tensor > nl || base_tensor < 1 || base_tensor > nl || net->tsize[base_tensor] != 1)
    return fail(XTB_ERR_ARG, "xtb_impala_train: bad head tensors");
  if (n_sample <= 0 || n_sample > net->max_batch || step_len < 2 || n_sample % step_len) return fail(XTB_ERR_ARG, "xtb_impala_train: bad sizes");
  const int adim = net->tsize[logit_tensor];
  if (adim > MAX_ADIM) return fail(XTB_ERR_ARG, "xtb_impala_train: action dim too large");
  StepKey key;
  memset(&key, 0, sizeof key);
  key.p[0] = net; key.p[1] = opt; key.p[2] = obs; key.p[3] = gather_idx; key.p[4] = bp_logits; key.p[5] = action; key.p[6] = done;
  key.p[7] = reward; key.p[8] = loss_out; key.p[9] = net->ws; key.p[10] = g_comm;
  key.i[0] = n_sample; key.i[1] = step_len; key.i[2] = logit_tensor; key.i[3] = base_tensor; key.i[4] = g_tc_mode; key.i[5] = 1;
  key.f[0] = gamma;
  return run_step_graph(key, use_graph, stream, [&](void* st) -> int {
    int rc = net_forward_impl(net, nullptr, obs, gather_idx, n_sample, st, 0u, (1u << logit_tensor) | (1u << base_tensor));
    if (rc) return rc;
    rc = xtb_vtrace_loss_grad(xtb_net_tensor(net, logit_tensor), xtb_net_tensor(net, base_tensor), bp_logits, action, done, reward,
                              n_sample / step_len, step_len, adim, gamma, xtb_net_tensor_grad(net, logit_tensor),
                              xtb_net_tensor_grad(net, base_tensor), nullptr, nullptr, loss_out, st);
    if (rc) return rc;
    int heads[2] = {logit_tensor, base_tensor};
    rc = net_backward_impl(net, obs, gather_idx, n_sample, heads, 2, st, 0u, true, 0u, 0u, g_comm);
    if (rc) return rc;
    return xtb_adam_step_net(opt, net, 1.f, st);
  });
}

// DQN.train (xt/algorithm/dqn/dqn.py:61-103) on a device replay ring: rows idx[0..n) of (obs, next_obs, action, reward,
// done[, disc]); target-network forward on s', optional double-DQN online forward on s', online forward on s, TD target +
// loss gradient, backward, clip + Adam.  qn_t / qn_o: scratch [n, adim] (qn_o NULL = plain DQN).  Reference mode:
// disc = NULL, huber_delta = 0.  With a communicator every rank holds n of world*n samples: inv_count = 1/(world*n*adim).
extern "C" int xtb_dqn_train(xtb_net* net, xtb_net* target, xtb_adam* opt, const void* obs, const void* next_obs,
                             const int32_t* idx, const int32_t* action, const float* reward, const uint8_t* done,
                             const float* disc, int n_sample, float gamma, float huber_delta, int q_tensor, float* qn_t,
                             float* qn_o, float* loss_out, int use_graph, void* stream) {
  if (!net || !target || !opt || !obs || !next_obs || !action || !reward || !done || !qn_t || !loss_out)
    return fail(XTB_ERR_ARG, "xtb_dqn_train: null pointer");
  if (!net->ws || !net->grads || !target->ws) return fail(XTB_ERR_STATE, "xtb_dqn_train: nets not bound");
  const int nl = (int)net->L.size();
  if (q_tensor < 1 || q_tensor > nl || (int)target->L.size() != nl) return fail(XTB_ERR_ARG, "xtb_dqn_train: bad head tensor");
  if (n_sample <= 0 || n_sample > net->max_batch || n_sample > target->max_batch) return fail(XTB_ERR_ARG, "xtb_dqn_train: bad batch");
  const int adim = net->tsize[q_tensor];
  StepKey key;
  memset(&key, 0, sizeof key);
  key.p[0] = net; key.p[1] = opt; key.p[2] = obs; key.p[3] = next_obs; key.p[4] = idx; key.p[5] = action; key.p[6] = reward;
  key.p[7] = done; key.p[8] = loss_out; key.p[9] = net->ws; key.p[10] = g_comm; key.p[11] = target;
  key.i[0] = n_sample; key.i[1] = q_tensor; key.i[2] = qn_o ? 1 : 0; key.i[3] = disc ? 1 : 0; key.i[4] = g_tc_mode; key.i[5] = 2;
  key.f[0] = gamma; key.f[1] = huber_delta;
  const float inv_world = g_comm ? 1.f / g_comm->world : 1.f;
  return run_step_graph(key, use_graph, stream, [&](void* st) -> int {
    const size_t qbytes = (size_t)n_sample * adim * sizeof(float);
    int rc = net_forward_impl(target, nullptr, next_obs, idx, n_sample, st, 0u, 1u << q_tensor);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(qn_t, xtb_net_tensor(target, q_tensor), qbytes, cudaMemcpyDeviceToDevice, S(st)));
    if (qn_o) {
      rc = net_forward_impl(net, nullptr, next_obs, idx, n_sample, st, 0u, 1u << q_tensor);
      if (rc) return rc;
      CUDA_TRY(cudaMemcpyAsync(qn_o, xtb_net_tensor(net, q_tensor), qbytes, cudaMemcpyDeviceToDevice, S(st)));
    }
    rc = net_forward_impl(net, nullptr, obs, idx, n_sample, st, 0u, 1u << q_tensor);
    if (rc) return rc;
    rc = xtb_dqn_td_loss_grad(xtb_net_tensor(net, q_tensor), qn_t, qn_o, idx, action, reward, done, disc, n_sample, adim, gamma,
                              huber_delta, inv_world / ((float)n_sample * adim), xtb_net_tensor_grad(net, q_tensor), nullptr, loss_out, st);
    if (rc) return rc;
    int heads[1] = {q_tensor};
    rc = net_backward_impl(net, obs, idx, n_sample, heads, 1, st, 0u, true, 0u, 0u, g_comm);
    if (rc) return rc;
    return xtb_adam_step_net(opt, net, 1.f, st);
  });
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
    if (it->first.net == obj || it->first.opt == obj || it->first.comm == obj) { cudaGraphExecDestroy(it->second.exec); it = g_graphs.erase(it); } else ++it;
  }
  for (auto it = g_infer_graphs.begin(); it != g_infer_graphs.end();) {
    if (it->first.net == obj) { cudaGraphExecDestroy(it->second.exec); it = g_infer_graphs.erase(it); } else ++it;
  }
  drop_step_graphs_of(obj);
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
    unsigned want = fuse ? ((1u << lpi.d.src) | (1u << lv.d.src)) : ((1u << pi_t) | (1u << v_t));
    int rc = net_forward_impl(net, nullptr, obs, step_idx ? step_idx + (long long)t * E : nullptr, E, stream, skip, want);
    if (rc) return rc;
    int32_t* a_t = action + (long long)t * E; float* lp_t = logp + (long long)t * E; float* v_o = value + (long long)t * E;
    if (fuse) {
      const float* hp = (const float*)(net->ws + net->out_off[lpi.d.src]);
      const float* hv = (const float*)(net->ws + net->out_off[lv.d.src]);
      const float *wp = net->params + lpi.w_off, *bp = net->params + lpi.b_off, *wv = net->params + lv.w_off, *bv = net->params + lv.b_off;
      int blocks = std::max(1, std::min(148, (E + 7) / 8));
      if (kpl <= 2) XLAUNCH((ppo_infer_heads_kernel<2, 8>), blocks, 256, 0, S(stream), hp, hv, wp, bp, wv, bv, E, lpi.K, adim, seed, offset_dev, t, a_t, lp_t, v_o, xtb_net_tensor(net, pi_t));
      else if (kpl <= 8 && adim <= 4) XLAUNCH((ppo_infer_heads_kernel<8, 4>), blocks, 256, 0, S(stream), hp, hv, wp, bp, wv, bv, E, lpi.K, adim, seed, offset_dev, t, a_t, lp_t, v_o, xtb_net_tensor(net, pi_t));
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
  if (!use_graph)
    return rollout_infer_launch(net, obs, step_idx, n_env, n_step, pi_tensor, v_tensor, seed, offset_dev, action, logp, value, stream);
  StreamScope sc;
  int src = sc.begin(stream, true);
  if (src) return src;
  InferKey key;
  memset(&key, 0, sizeof key);
  key.net = net; key.obs = obs; key.idx = step_idx; key.act = action; key.logp = logp; key.val = value; key.ctr = offset_dev;
  key.ws = net->ws; key.seed = seed; key.e = n_env; key.t = n_step; key.pi_t = pi_tensor; key.v_t = v_tensor; key.tc = g_tc_mode;
  auto it = g_infer_graphs.find(key);
  if (it == g_infer_graphs.end()) {
    cudaStream_t st = sc.st;
    long long before = g_launches.load();
    CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = rollout_infer_launch(net, obs, step_idx, n_env, n_step, pi_tensor, v_tensor, seed, offset_dev, action, logp, value, (void*)st);
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
  CUDA_TRY(cudaGraphLaunch(it->second.exec, sc.st));
  g_launches.fetch_add(it->second.kernels, std::memory_order_relaxed);
  g_graph_replays.fetch_add(1, std::memory_order_relaxed);
  return sc.end();
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
extern "C" int xtb_actor_predict_host(xtb_net* net, const void* obs_host, size_t obs_bytes, void* obs_dev, int n_env,
                                      int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                                      float* out_dev, float* out_host, float* logits_host, int use_graph, void* stream) {
  if (!net || !obs_host || !obs_dev || !out_dev || !out_host) return fail(XTB_ERR_ARG, "xtb_actor_predict_host: null pointer");
  StreamScope sc;
  int src = sc.begin(stream, use_graph != 0);
  if (src) return src;
  void* st = (void*)sc.st;
  CUDA_TRY(xtb::Stager::instance().stage_h2d(obs_dev, obs_host, obs_bytes, sc.st));
  int rc = xtb_ppo_rollout_infer(net, obs_dev, nullptr, n_env, 1, pi_tensor, v_tensor, seed, offset_dev,
                                 reinterpret_cast<int32_t*>(out_dev), out_dev + n_env, out_dev + 2 * (size_t)n_env, use_graph, st);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_host, out_dev, sizeof(float) * 3 * (size_t)n_env, cudaMemcpyDeviceToHost, sc.st));
  if (logits_host)
    CUDA_TRY(cudaMemcpyAsync(logits_host, xtb_net_tensor(net, pi_tensor), sizeof(float) * (size_t)n_env * net->tsize[pi_tensor],
                             cudaMemcpyDeviceToHost, sc.st));
  CUDA_TRY(cudaStreamSynchronize(sc.st));
  return sc.end();
}
extern "C" int xtb_ppo_predict_host(xtb_net* net, const void* obs_host, size_t obs_bytes, void* obs_dev, int n_env,
                                    int pi_tensor, int v_tensor, uint64_t seed, unsigned long long* offset_dev,
                                    float* out_dev, float* out_host, int use_graph, void* stream) {
  return xtb_actor_predict_host(net, obs_host, obs_bytes, obs_dev, n_env, pi_tensor, v_tensor, seed, offset_dev, out_dev, out_host,
                                nullptr, use_graph, stream);
}
extern "C" int xtb_copy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, S(stream)));
  return XTB_OK;
}
extern "C" int xtb_stream_sync(void* stream) {
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  return XTB_OK;
}
