// mvb_api.cu -- the C-ABI of libmvb.so (include/mvb.h): argument checking, workspace carving and
// the kernel sequences of the ridge hot path.  No allocation, no synchronisation, no host fallback.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mvb.h"
#include "mvb_kernels.cuh"
#include "tc_gemm.cuh"
#include "toeplitz_gram.cuh"
#include "gram_recurrence.cuh"
#include "jacobi_block.cuh"
#include "lanczos_band.cuh"
#include "band_loo.cuh"
#include "sweep_chain.cuh"
#include "rf_kernels.cuh"
#include "dense_batched.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define MVB_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return fail(MVB_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define MVB_LAUNCHED()                                                                   \
  do {                                                                                   \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess)                                                              \
      return fail(MVB_E_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- measurement hook (include/mvb.h) ------------------------------------------------------------------
enum ProfClass { PC_GRAM = 0, PC_XTY, PC_SMALL, PC_LEVERAGE_Z, PC_RESIDUAL, PC_PREDICT, PC_EIG_GRAM, PC_EIG_APPLY,
                 PC_EIG_REFRESH, PC_OTHER, PC_COUNT };
const char* kProfNames[PC_COUNT] = {"gram X^T X (tensor-core part: first block rows, or the dense Toeplitz kernel)", "X^T y", "Q b and coef = Q^T x (dense)", "Z = X Q^T (leverages)",
                                    "R = Z S^T (LOO fitted values, reduced basis)", "predict X coef", "eig: pair Gram", "eig: rotate rows",
                                    "band reduction (block Lanczos) / eig refresh", "block substitution sweep (forward / backward)"};
constexpr int kProfMaxTimed = 32768;   // per class: enough to time every contraction launch of a multi-step bench run
struct ProfState {
  bool on = false;
  std::mutex mu;
  cudaEvent_t ev[PC_COUNT][kProfMaxTimed][2];
  bool created[PC_COUNT][kProfMaxTimed] = {};
  int n_timed[PC_COUNT] = {};
  int64_t n_all[PC_COUNT] = {};
  double flops_timed[PC_COUNT] = {}, flops_all[PC_COUNT] = {};
} g_prof;
thread_local int g_prof_class = PC_OTHER;
struct ProfScope {
  int prev;
  explicit ProfScope(int c) : prev(g_prof_class) { g_prof_class = c; }
  ~ProfScope() { g_prof_class = prev; }
};

// returns slot (>=0) if this launch is timed; records the start event
int prof_begin(double flops, cudaStream_t st) {
  if (!g_prof.on) return -1;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  const int c = g_prof_class;
  g_prof.n_all[c]++;
  g_prof.flops_all[c] += flops;
  if (g_prof.n_timed[c] >= kProfMaxTimed) return -1;
  const int slot = g_prof.n_timed[c]++;
  if (!g_prof.created[c][slot]) {
    cudaEventCreate(&g_prof.ev[c][slot][0]);
    cudaEventCreate(&g_prof.ev[c][slot][1]);
    g_prof.created[c][slot] = true;
  }
  g_prof.flops_timed[c] += flops;
  cudaEventRecord(g_prof.ev[c][slot][0], st);
  return slot;
}
void prof_end(int slot, cudaStream_t st) {
  if (slot < 0) return;
  cudaEventRecord(g_prof.ev[g_prof_class][slot][1], st);
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int grid_for(int64_t total, int threads = 256, int cap = 148 * 16) {
  int64_t g = (total + threads - 1) / threads;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

bool force_simt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MVB_FORCE_SIMT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// bump allocator over the caller's workspace
struct Arena {
  char* base; size_t size; size_t off = 0; bool dry;
  Arena(void* b, size_t s) : base((char*)b), size(s), dry(b == nullptr) {}
  template <typename U> U* take(size_t count) {
    off = align_up(off);
    U* p = dry ? nullptr : (U*)(base + off);
    off += count * sizeof(U);
    return p;
  }
  bool ok() const { return dry || off <= size; }
};

// ---- contraction dispatch ---------------------------------------------------------------------------
struct SplitPlan { int splits; size_t ws_elems; };

SplitPlan plan_contract(int M, int N, int K, int dtype, int symmetric) {
  SplitPlan pl{1, 0};
  const bool tc = (dtype == MVB_F32) && !force_simt();
  const int bm = tc ? 128 : 64;
  const int bn = tc ? (N > 128 ? 256 : 128) : 64;
  const int bk = tc ? 32 * mvb::tc::DRAIN : 64;
  int64_t tiles = (int64_t)cdiv(M, bm) * cdiv(N, bn);
  if (symmetric) tiles = tiles / 2 + cdiv(M, bm);
  const int kblocks = cdiv(K, bk);
  const int target = tc ? 148 : 296;
  int s = 1;
  if (tc) {
    s = mvb::tc_choose_splits(tiles, cdiv(K, 32), 64, (int64_t)M * N);
  } else {
    if (tiles < target) s = (int)((target + tiles - 1) / tiles);
    if (s > kblocks) s = kblocks;
    if (s > 64) s = 64;
  }
  if (s < 1) s = 1;
  pl.splits = s;
  pl.ws_elems = s > 1 ? (size_t)s * M * N : 0;
  return pl;
}

template <typename T>
int run_contract(const mvb_operand& A, const mvb_operand& B, int M, int N, int K, T* C, int64_t ldc, int symmetric,
                 void* ws, size_t ws_bytes, cudaStream_t st, int c_block_cols = 0, int64_t c_block_stride = 0,
                 void* planes = nullptr, size_t planes_bytes = 0) {
  if (M <= 0 || N <= 0) return MVB_OK;
  constexpr int dtype = sizeof(T) == 4 ? MVB_F32 : MVB_F64;
  const SplitPlan pl = plan_contract(M, N, K, dtype, symmetric);
  T* out = C;
  int64_t ld = ldc, stride = 0;
  if (pl.splits > 1) {
    if (ws_bytes < pl.ws_elems * sizeof(T)) return fail(MVB_E_WORKSPACE, "contract: workspace %zu < %zu", ws_bytes, pl.ws_elems * sizeof(T));
    out = (T*)ws; ld = N; stride = (int64_t)M * N;
  }
  const bool tc = (dtype == MVB_F32) && !force_simt();
  if (tc) {
    mvb::TcGemmParams p = mvb::tc_gemm_defaults();
    p.A = {(const float*)A.base, A.roff, A.koff, A.rows, A.contig};
    p.B = {(const float*)B.base, B.roff, B.koff, B.rows, B.contig};
    p.M = M; p.N = N; p.K = K; p.C = (float*)out; p.ldc = ld; p.c_split_stride = stride; p.symmetric = symmetric; p.dbg = 0;
    if (pl.splits == 1) { p.c_block_cols = c_block_cols; p.c_block_stride = c_block_stride; }   // partials stay row-major
    // a B operand that is larger than L2 (dense basis / coefficient matrices against the small implicit design) is
    // streamed once per wave of M tiles instead of once per few M tiles
    p.raster_m_fast = (!symmetric && (int64_t)N * K * 4 > ((int64_t)64 << 20) && M >= 148 * 128 && N <= 65535 * 128) ? 1 : 0;
    // Large products against a dense B: B is split into TF32 hi / lo planes once (shared-memory image per 256 x 32 tile)
    // and every stage's B half then arrives by one TMA bulk copy instead of LDG -> split -> STS in the 512 threads,
    // which is the bound of this kernel (profiles/r01_tc_gemm2_notes.md); worth one extra pass over B when M is large.
    static const bool planes_on = [] { const char* e = getenv("MVB_TC_PLANES"); return !e || atoi(e) != 0; }();
    // (symmetric products never take planes: the Gram of a lag-expanded design has its own kernel, toeplitz_gram.cuh)
    if (planes && planes_on && !symmetric && N > 128 && M >= 1024 &&
        (double)M * N * K >= 4e9 && planes_bytes >= mvb::tc_planes_bytes(N, K)) {
      MVB_CUDA(mvb::tc_presplit_launch(p.B, K, (uint8_t*)planes, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
      p.b_planes = (const uint8_t*)planes;
      p.b_planes_nkb = (K + mvb::tc::BK - 1) / mvb::tc::BK;
      static const bool reuse_hint = [] { const char* e = getenv("MVB_TC_REUSE_HINTS"); return !e || atoi(e) != 0; }();
      if (reuse_hint && p.raster_m_fast) p.b_stream_once = 2;      // B tile re-read by every M tile of the wave: keep it in L2, let A stream
    }
    const int slot = prof_begin(2.0 * M * (double)N * K * (symmetric ? 0.5 : 1.0), st);
    MVB_CUDA(mvb::tc_gemm_launch(p, pl.splits, N > 128 ? 256 : 128, st));
    prof_end(slot, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    mvb::GOp<T> a{(const T*)A.base, A.roff, A.koff, A.rows}, b{(const T*)B.base, B.roff, B.koff, B.rows};
    dim3 grid(cdiv(N, 64), cdiv(M, 64), pl.splits);
    if (c_block_cols > 0 && pl.splits == 1) return fail(MVB_E_UNSUPPORTED, "contract: column-blocked output needs the tensor-core path");
    mvb::simt_contract_kernel<T><<<grid, 256, 0, st>>>(a, b, M, N, K, out, ld, stride, symmetric);
    MVB_LAUNCHED();
  }
  if (pl.splits > 1) {
    mvb::reduce_splits_kernel<T><<<grid_for((int64_t)M * N), 256, 0, st>>>(out, pl.splits, stride, M, N, ld, C, ldc, c_block_cols,
                                                                              c_block_stride);
    MVB_LAUNCHED();
  }
  return MVB_OK;
}

int dense_mode(int64_t ld, int K, const void* base, size_t elem) {
  return (elem == 4 && ld % 4 == 0 && K % 4 == 0 && ((uintptr_t)base & 15) == 0) ? 2 : 1;
}

bool design_ok(const mvb_design* d) {
  return d && d->xp && d->n_time > 0 && d->n_lag > 0 && d->n_trials > 0 && d->n_feat > 0 &&
         (d->dtype == MVB_F32 || d->dtype == MVB_F64);
}

// ---- stats ------------------------------------------------------------------------------------------
struct StatsWs {
  int64_t *row_off, *col_off, *yrow_off, *ycol_off;
  double *partial, *mu64, *ybar64;
  void* contract_ws; size_t contract_bytes;
  void* planes; size_t planes_bytes;      // pre-split TF32 planes of y (the dense B operand of X^T y), large fp32 problems only
  float* pack; size_t pack_bytes;         // TF32 planes of the four shifted copies of the training series (toeplitz_gram.cuh)
  float* B0; int64_t* lag0_off;           // first block rows G[(c,0),:] and the offsets of the lag-0 columns (gram_recurrence.cuh)
  int parts;
};

// Default route of the Gram of a lag-expanded fp32 design: its displacement structure (gram_recurrence.cuh), 1 % of the dense
// contraction's flops.  MVB_GRAM_RECURRENCE=0 takes the dense routes below (the implicit-Toeplitz tensor-core kernel when its shape
// allows, else the generic gather kernel).
bool use_gram_recurrence(const mvb_design* d) {
  const char* e = getenv("MVB_GRAM_RECURRENCE");
  if (e && atoi(e) == 0) return false;
  return d->dtype == MVB_F32 && !force_simt() && mvb::gr::usable(d->n_lag);
}

// The Gram of a lag-expanded design runs on the implicit-Toeplitz kernel (toeplitz_gram.cuh) whenever its shape allows:
// fp32, tensor cores, whole 8-sample MMA steps per trial, enough tiles to fill the SMs.  MVB_TOEPLITZ_GRAM=0 keeps the
// generic gather kernel (A/B measurements), =2 takes it for every shape it can compute, however few tiles (tests).
bool use_toeplitz_gram(const mvb_design* d) {
  const char* e = getenv("MVB_TOEPLITZ_GRAM");
  const int mode = e ? atoi(e) : 1;
  if (mode == 0 || d->dtype != MVB_F32 || force_simt() || use_gram_recurrence(d)) return false;
  return mvb::tg::usable(d->n_time, d->n_lag, d->n_feat, mode == 2);
}

template <typename T>
size_t carve_stats(Arena& ar, const mvb_design* d, int F, StatsWs& w) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  w.parts = (int)std::min<int64_t>(64, std::max<int64_t>(1, n / 256));
  w.row_off = ar.take<int64_t>(n);
  w.col_off = ar.take<int64_t>(p);
  w.yrow_off = ar.take<int64_t>(n);
  w.ycol_off = ar.take<int64_t>(F);
  w.partial = ar.take<double>((size_t)w.parts * (p + F));
  w.mu64 = ar.take<double>(p);
  w.ybar64 = ar.take<double>(F);
  constexpr int dt = sizeof(T) == 4 ? MVB_F32 : MVB_F64;
  const bool rec = sizeof(T) == 4 && use_gram_recurrence(d);
  w.B0 = rec ? ar.take<float>((size_t)d->n_feat * p) : nullptr;
  w.lag0_off = rec ? ar.take<int64_t>(d->n_feat) : nullptr;
  size_t cb = std::max(rec ? plan_contract(d->n_feat, (int)p, (int)n, dt, 0).ws_elems : plan_contract((int)p, (int)p, (int)n, dt, 1).ws_elems,
                       plan_contract((int)p, F, (int)n, dt, 0).ws_elems) * sizeof(T);
  w.contract_bytes = cb;
  w.contract_ws = ar.take<char>(cb);
  w.planes_bytes = (sizeof(T) == 4 && p >= 1024 && !force_simt()) ? mvb::tc_planes_bytes(F, (int)n) : 0;
  w.planes = w.planes_bytes ? ar.take<char>(w.planes_bytes) : nullptr;
  w.pack_bytes = use_toeplitz_gram(d) ? mvb::tg::pack_bytes(d->n_trials, d->n_feat, d->n_time + d->n_lag - 1) : 0;
  w.pack = w.pack_bytes ? (float*)ar.take<char>(w.pack_bytes) : nullptr;
  return ar.off;
}

template <typename T>
int trf_stats_impl(const mvb_design* d, const T* y, int F, int fit_intercept, T* G, T* XtY, T* mu, T* ybar, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  Arena ar(ws, ws_bytes);
  StatsWs w;
  carve_stats<T>(ar, d, F, w);
  if (!ar.ok()) return fail(MVB_E_WORKSPACE, "trf_stats: workspace %zu < %zu", ws_bytes, ar.off);
  const int Tn = d->n_time;
  mvb::design_offsets_kernel<<<grid_for(n + p), 256, 0, st>>>(d->trial_idx, d->n_trials, Tn, d->trial_stride, d->feat_idx,
                                                            d->n_feat, d->n_lag, d->feat_stride, (int64_t)F * Tn,
                                                            w.row_off, w.col_off, w.yrow_off);
  MVB_LAUNCHED();
  mvb::iota_offsets_kernel<<<grid_for(F), 256, 0, st>>>(w.ycol_off, F, Tn);
  MVB_LAUNCHED();
  // column / target sums -> means
  {
    dim3 gx(cdiv(p, 128), w.parts), gy(cdiv(F, 128), w.parts);
    mvb::col_sums_kernel<T><<<gx, 128, 0, st>>>((const T*)d->xp, w.row_off, w.col_off, n, p, w.partial);
    MVB_LAUNCHED();
    mvb::col_sums_kernel<T><<<gy, 128, 0, st>>>(y, w.yrow_off, w.ycol_off, n, F, w.partial + (size_t)w.parts * p);
    MVB_LAUNCHED();
    mvb::sums_finish_kernel<T><<<cdiv(p, 256), 256, 0, st>>>(w.partial, w.parts, p, 1.0 / (double)n, fit_intercept, mu, w.mu64);
    MVB_LAUNCHED();
    mvb::sums_finish_kernel<T><<<cdiv(F, 256), 256, 0, st>>>(w.partial + (size_t)w.parts * p, w.parts, F, 1.0 / (double)n,
                                                             fit_intercept, ybar, w.ybar64);
    MVB_LAUNCHED();
  }
  // design columns as tile rows: consecutive lags are consecutive samples (mode 0), and so are consecutive k (t)
  mvb_operand xc{d->xp, w.col_off, w.row_off, (int)p, 0};
  const bool y_vec = sizeof(T) == 4 && Tn % 4 == 0 && ((uintptr_t)y & 15) == 0;
  mvb_operand yc{y, w.ycol_off, w.yrow_off, F, y_vec ? 2 : 1};
  int rc;
  if constexpr (sizeof(T) == 4) {
    if (w.B0) {
      // X^T X from its displacement structure: the first row of every (c, c') block by one skinny tensor-core contraction over
      // the implicit design (the C lag-0 columns against all p columns), the rest of each block by the diagonal recurrence
      mvb::stride_pick_kernel<<<grid_for(d->n_feat), 256, 0, st>>>(w.col_off, d->n_lag, d->n_feat, w.lag0_off);
      MVB_LAUNCHED();
      mvb_operand x0{d->xp, w.lag0_off, w.row_off, d->n_feat, 1};          // rows = lag-0 columns, k = (trial, t) contiguous in t
      { ProfScope ps(PC_GRAM); rc = run_contract<float>(x0, xc, d->n_feat, (int)p, (int)n, w.B0, p, 0, w.contract_ws, w.contract_bytes, st); }
      if (rc) return rc;
      MVB_CUDA(mvb::gr::launch((const float*)d->xp, d->trial_idx, d->n_trials, d->trial_stride, d->feat_idx, d->n_feat, d->feat_stride, d->n_lag,
                               Tn, w.B0, (float*)G, p, st));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (w.pack_bytes) {
      // X^T X straight from the padded stimulus: the training series are split into TF32 planes once (8 shifted / split
      // copies, 130 MB at configs[1]) and the kernel's operands are windows of them -- nothing of the size of the lag
      // expansion (timedelayed.py:415-423: 1.98 GB) exists at any point
      namespace tg = mvb::tg;
      const int Tp = Tn + d->n_lag - 1, Tpp = tg::pack_pitch(Tp);
      tg::pack_kernel<<<148 * 8, 256, 0, st>>>((const float*)d->xp, d->trial_idx, d->n_trials, d->trial_stride, d->feat_idx, d->n_feat,
                                               d->feat_stride, Tp, Tpp, w.pack);
      MVB_LAUNCHED();
      tg::Params tp{};
      tp.pack = w.pack; tp.L = (int64_t)d->n_trials * d->n_feat * Tpp; tp.n_trials = d->n_trials; tp.n_feat = d->n_feat; tp.Tpp = Tpp;
      tp.nblk_f = (d->n_lag + 31) / 32; tp.n_vblk = d->n_feat * tp.nblk_f; tp.W = d->n_lag; tp.T = Tn;
      tg::chunk_plan(Tn, tp.n_chunks, tp.chunk_len);
      tp.G = (float*)G; tp.ldg = p;
      ProfScope ps(PC_GRAM);
      const int slot = prof_begin((double)n * (double)p * (double)p, st);
      MVB_CUDA(tg::launch(tp, st));
      prof_end(slot, st);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      rc = MVB_OK;
    } else {
      ProfScope ps(PC_GRAM); rc = run_contract<T>(xc, xc, (int)p, (int)p, (int)n, G, p, 1, w.contract_ws, w.contract_bytes, st);
    }
  } else {
    ProfScope ps(PC_GRAM); rc = run_contract<T>(xc, xc, (int)p, (int)p, (int)n, G, p, 1, w.contract_ws, w.contract_bytes, st);
  }
  if (rc) return rc;
  { ProfScope ps(PC_XTY); rc = run_contract<T>(xc, yc, (int)p, F, (int)n, XtY, F, 0, w.contract_ws, w.contract_bytes, st, 0, 0, w.planes, w.planes_bytes); }
  if (rc) return rc;
  if (fit_intercept) {
    mvb::centre_moments_kernel<T><<<grid_for(p * p + p * F), 256, 0, st>>>(G, p, XtY, F, w.mu64, w.ybar64, (double)n);
    MVB_LAUNCHED();
  }
  return MVB_OK;
}

// ---- eigensolver -------------------------------------------------------------------------------------
constexpr int kMaxSweeps = 40;
constexpr int kBlockJacobiMinP = 160;   // fp32 problems at least this large go to the tensor-core block Jacobi

inline bool use_block_jacobi(int p, int dtype) { return dtype == MVB_F32 && p >= kBlockJacobiMinP && !force_simt(); }

size_t eigh_ws_bytes(int batch, int p, int dtype) {
  size_t b = (size_t)batch * kMaxSweeps * sizeof(unsigned int) + 256;
  if (use_block_jacobi(p, dtype)) b = std::max(b, mvb::bj::make_plan(p).bytes);
  return b;
}

template <typename T>
int eigh_impl(T* A, T* Vt, T* lam, int batch, int p, void* ws, size_t ws_bytes, cudaStream_t st) {
  constexpr int dtype = sizeof(T) == 4 ? MVB_F32 : MVB_F64;
  const size_t need = eigh_ws_bytes(batch, p, dtype);
  if (ws_bytes < need) return fail(MVB_E_WORKSPACE, "eigh: workspace %zu < %zu", ws_bytes, need);
  if (use_block_jacobi(p, dtype)) {
    for (int b = 0; b < batch; ++b) {
      cudaError_t err = cudaSuccess;
      mvb::bj::Hooks hooks{[](int cls, double flops, cudaStream_t s_) { g_prof_class = PC_EIG_GRAM + cls; return prof_begin(flops, s_); },
                           [](int slot, cudaStream_t s_) { prof_end(slot, s_); g_prof_class = PC_OTHER; }};
      const int n = mvb::bj::solve((const float*)A + (size_t)b * p * p, (float*)Vt + (size_t)b * p * p, (float*)lam + (size_t)b * p, p, ws, st, &err, &hooks);
      if (n < 0) return fail(MVB_E_CUDA, "block Jacobi eigensolver failed: %s", cudaGetErrorString(err));
      g_launches.fetch_add(n, std::memory_order_relaxed);
    }
    return MVB_OK;
  }
  MVB_CUDA(cudaMemsetAsync(ws, 0, (size_t)batch * kMaxSweeps * sizeof(unsigned int), st));
  int cluster = 1;
  if (p > 96) cluster = 8;
  int threads = p <= 64 ? 256 : 1024;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(batch * cluster);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const double eps = sizeof(T) == 4 ? 5.96e-8 : 1.11e-16;
  const double tol = eps * sqrt((double)(p < 4 ? 4 : p));
  MVB_CUDA(cudaLaunchKernelEx(&cfg, mvb::jacobi_rows_kernel<T>, A, Vt, lam, p, kMaxSweeps, tol, (unsigned int*)ws, cluster));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MVB_OK;
}

// ---- ridge solve ---------------------------------------------------------------------------------------
struct SolveWs {
  int64_t *row_off, *col_off, *yrow_off, *iota_p, *iota_1, *iota_F;
  void *Vt, *lam, *w, *Vtb, *theta, *beta, *ZR, *d;
  double *mu64, *ybar64, *cz, *mb, *loss_partial, *loss64;
  void* eig_ws; size_t eig_bytes;
  void* contract_ws; size_t contract_bytes;
  int slabs;
};

template <typename T>
void carve_solve(Arena& ar, const mvb_design* d, int F, int A, SolveWs& w) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  const int64_t AF = (int64_t)A * F;
  constexpr int dt = sizeof(T) == 4 ? MVB_F32 : MVB_F64;
  w.slabs = (int)std::min<int64_t>(128, std::max<int64_t>(1, n / 128));
  w.row_off = ar.take<int64_t>(n);
  w.col_off = ar.take<int64_t>(p);
  w.yrow_off = ar.take<int64_t>(n);
  w.iota_p = ar.take<int64_t>(std::max<int64_t>(p, AF));
  w.iota_1 = ar.take<int64_t>(std::max<int64_t>(p, AF));
  w.iota_F = ar.take<int64_t>(p);
  w.Vt = ar.take<T>(p * p);
  w.lam = ar.take<T>(p);
  w.w = ar.take<T>((size_t)A * p);
  w.Vtb = ar.take<T>(p * F);
  w.theta = ar.take<T>(AF * p);
  w.beta = ar.take<T>(AF * p);
  w.ZR = ar.take<T>(n * std::max<int64_t>(p, AF));
  w.d = ar.take<T>(n * A);
  w.mu64 = ar.take<double>(p);
  w.ybar64 = ar.take<double>(F);
  w.cz = ar.take<double>(p);
  w.mb = ar.take<double>(AF);
  w.loss_partial = ar.take<double>((size_t)w.slabs * AF);
  w.loss64 = ar.take<double>(AF);
  w.eig_bytes = eigh_ws_bytes(1, (int)p, dt);
  w.eig_ws = ar.take<char>(w.eig_bytes);
  size_t cb = 0;
  cb = std::max(cb, plan_contract((int)p, F, (int)p, dt, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)AF, (int)p, (int)p, dt, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)n, (int)p, (int)p, dt, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)n, (int)AF, (int)p, dt, 0).ws_elems);
  w.contract_bytes = cb * sizeof(T);
  w.contract_ws = ar.take<char>(w.contract_bytes);
}

template <typename T>
__global__ void to_double_kernel(const T* __restrict__ in, int64_t n, double* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}


// ---- band path: block-tridiagonal reduction + block-Cholesky LOO sweep (fp32, large p) ---------------------
int band_min_p() {
  const char* e = getenv("MVB_BAND_MIN_P");
  int v = e ? atoi(e) : 128;
  if (v < 128) v = 128;
  return v;
}
inline bool use_band(int64_t p, int dtype) { return dtype == MVB_F32 && p >= band_min_p() && !force_simt(); }

struct BandWs {
  int64_t *row_off, *col_off, *yrow_off, *iota, *rowP, *tab128, *tab256, *kofs, *rowUb, *iota_F, *iota_p, *zkoff;
  float *R, *Ssel;             // fitted values of all (alpha, target) pairs, n x AF; substitution result of the selected alpha(s), F x P
  float *Qt, *Tdiag, *Tsub, *Minv, *Nmat, *Pb, *MN, *Z, *h, *U0, *U1, *C1, *C2, *Ct, *Ub, *Xb, *beta, *dd;
  double *mu64, *ybar64, *cz, *mb, *loss_partial, *loss64;
  void* lanczos_ws; void* contract_ws; size_t contract_bytes;
  void* planes; size_t planes_bytes;      // pre-split TF32 planes of the dense B operand of the large products
  int slabs, P, nb;
};

void carve_band(Arena& ar, const mvb_design* d, int F, int A, BandWs& w) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  const int64_t AF = (int64_t)A * F;
  const mvb::lb::Plan lp = mvb::lb::make_plan((int)p);
  const int64_t P = lp.P, nb = lp.nblk;
  w.P = (int)P; w.nb = (int)nb;
  w.slabs = (int)std::min<int64_t>(128, std::max<int64_t>(1, n / 128));
  const int64_t rows_max = std::max<int64_t>(n, F);
  w.row_off = ar.take<int64_t>(n);
  w.col_off = ar.take<int64_t>(p);
  w.yrow_off = ar.take<int64_t>(n);
  w.iota = ar.take<int64_t>(std::max<int64_t>(P, AF) + 256);
  w.rowP = ar.take<int64_t>(std::max<int64_t>(std::max<int64_t>(P, n), AF));
  w.tab128 = ar.take<int64_t>(std::max<int64_t>((int64_t)A * nb * 128, (int64_t)A * rows_max));
  w.tab256 = ar.take<int64_t>((int64_t)A * nb * 128);
  w.kofs = ar.take<int64_t>((int64_t)nb * A * 256);
  w.rowUb = ar.take<int64_t>(AF);
  w.iota_F = ar.take<int64_t>(p);
  w.iota_p = ar.take<int64_t>(std::max<int64_t>(p, AF));
  w.Qt = ar.take<float>(P * P);
  w.Tdiag = ar.take<float>(nb * 128 * 128);
  w.Tsub = ar.take<float>(nb * 128 * 128);
  w.Minv = ar.take<float>((int64_t)A * nb * 128 * 128);
  w.Nmat = ar.take<float>((int64_t)A * nb * 128 * 128);
  w.Pb = ar.take<float>((int64_t)A * nb * 128 * 128);
  w.MN = ar.take<float>((int64_t)A * nb * 128 * 256);
  w.Z = ar.take<float>(n * P);                               // column-blocked [nb][n][128], centred: the sweep's and the fitted values' operand
  w.R = ar.take<float>(n * AF);
  w.Ssel = ar.take<float>((int64_t)F * P);
  w.zkoff = ar.take<int64_t>(P);
  w.h = ar.take<float>((int64_t)mvb::TC_ROWSQ_PARTS * A * n);   // column parts of sum_j |u_j|^2
  w.U0 = ar.take<float>((int64_t)A * n * 128);
  w.U1 = ar.take<float>((int64_t)A * n * 128);
  w.C1 = ar.take<float>((int64_t)A * F * 128);
  w.C2 = ar.take<float>((int64_t)A * F * 128);
  w.Ct = ar.take<float>((int64_t)F * P);
  w.Ub = ar.take<float>(AF * P);
  w.Xb = ar.take<float>(AF * P);
  w.dd = ar.take<float>(n * A);
  w.mu64 = ar.take<double>(P);
  w.ybar64 = ar.take<double>(F);
  w.cz = ar.take<double>(P);
  w.mb = ar.take<double>(AF);
  w.loss_partial = ar.take<double>((size_t)w.slabs * AF);
  w.loss64 = ar.take<double>(AF);
  w.lanczos_ws = ar.take<char>(lp.bytes);
  size_t cb = 0;
  cb = std::max(cb, plan_contract((int)F, (int)P, (int)p, MVB_F32, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)AF, (int)p, (int)P, MVB_F32, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)n, (int)P, (int)p, MVB_F32, 0).ws_elems);
  cb = std::max(cb, plan_contract((int)n, (int)AF, (int)P, MVB_F32, 0).ws_elems);
  cb = std::max(cb, plan_contract(F, (int)p, (int)P, MVB_F32, 0).ws_elems);
  w.contract_bytes = cb * 4;
  w.contract_ws = ar.take<char>(w.contract_bytes);
  w.planes_bytes = std::max(std::max(mvb::tc_planes_bytes((int)P, (int)p), mvb::tc_planes_bytes((int)p, (int)P)),
                            mvb::tc_planes_bytes((int)AF, (int)P));
  w.planes = ar.take<char>(w.planes_bytes);
}

// column-blocked Z [nb][rows][128]: Z[blk][r][c] -= cz[blk * 128 + c]
__global__ void sub_cols_blocked_kernel(float* __restrict__ Z, int64_t rows, int64_t nb, const double* __restrict__ c) {
  const int64_t total = nb * rows * 128;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t blk = e / (rows * 128);
    Z[e] = (float)((double)Z[e] - c[blk * 128 + (e & 127)]);
  }
}
// k offsets (relative to Zblk) of the fused forward substitution operand [ z_j | u_{j-1}(alpha) ], table [nb][A][256]:
// u_{j-1} lives in U1 for odd j and in U0 for even j (ping-pong); off_u0 / off_u1 = element distance of U0 / U1 from Zblk
__global__ void sweep_koff_kernel(int64_t* __restrict__ kofs, int nb, int A, int64_t n, int64_t off_u0, int64_t off_u1) {
  const int64_t total = (int64_t)nb * A * 256;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e & 255);
    const int64_t ja = e >> 8;
    const int64_t j = ja / A, a = ja - j * A;
    if (k < 128) kofs[e] = j * n * 128 + k;
    else kofs[e] = ((j & 1) ? off_u0 : off_u1) + a * n * 128 + (k - 128);
  }
}
__global__ void pad_double_kernel(const float* __restrict__ in, int64_t n_in, int64_t n_out, double* out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_out; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = e < n_in ? (double)in[e] : 0.0;
}
// d[i][a] = 1 - c0 - h[a][i]
__global__ void leverage_to_d_kernel(const float* __restrict__ h, int64_t n, int A, float c0, float* d, int halves) {
  const int64_t total = n * A;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / A; const int a = (int)(e - i * A);
    float hv = 0.f;
    for (int part = 0; part < halves; ++part) hv += h[((int64_t)part * A + a) * n + i];
    d[e] = 1.f - c0 - hv;
  }
}

// k offsets of the column-blocked Z ([P / 128][n][128]) as a contraction operand: k -> (k / 128) * n * 128 + k % 128
__global__ void zblock_koff_kernel(int64_t* __restrict__ koff, int64_t P, int64_t n) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (int64_t)gridDim.x * blockDim.x)
    koff[k] = (k >> 7) * n * 128 + (k & 127);
}
// out[f][:] = X[(best_f, f)][:]   (X: [A][F][P])
__global__ void gather_rows_kernel(const float* __restrict__ X, int F, int64_t P, const int32_t* __restrict__ best, int per_target,
                                   float* __restrict__ out) {
  const int64_t total = (int64_t)F * P;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(e / P);
    const int a = per_target ? best[f] : best[0];
    out[e] = X[((int64_t)a * F + f) * P + (e - (int64_t)f * P)];
  }
}
__global__ void intercept_kernel(const double* __restrict__ ybar, const double* __restrict__ mb, int F, int fit_intercept, float* intercept) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < F) intercept[f] = fit_intercept ? (float)(ybar[f] - mb[f]) : 0.f;
}

// helper stream for work that is independent of the caller's stream for a while (forked / joined with events): one per
// (device, caller's stream), created on first use with that device current.  Per caller's stream, because independent fits are
// enqueued on several stream lanes at once (mvpy_b200/_streams.py): on one shared helper stream the forked work of lane B would
// queue behind lane A's and B's join would wait for both.
cudaStream_t side_stream(cudaStream_t caller) {
  static std::map<std::pair<int, cudaStream_t>, cudaStream_t> streams;
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(dev, caller);
  auto it = streams.find(key);
  if (it != streams.end()) return it->second;
  if (streams.size() >= 256) {                                   // callers that churn through streams: share the first helper of the device
    for (auto& kv : streams)
      if (kv.first.first == dev) return kv.second;
  }
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  streams[key] = s;
  return s;
}
// Joins forked side-stream work back into the caller's stream on every exit path (an error return between fork and join must
// not leave the side stream writing into a workspace the caller is about to free, nor an open capture un-joined).
struct SideJoin {
  cudaStream_t st; cudaEvent_t ev = nullptr; bool armed = false;
  explicit SideJoin(cudaStream_t s) : st(s) {}
  void arm(cudaEvent_t e) { ev = e; armed = true; }
  cudaError_t join() { if (!armed) return cudaSuccess; armed = false; return cudaStreamWaitEvent(st, ev, 0); }
  ~SideJoin() { if (armed) cudaStreamWaitEvent(st, ev, 0); }
};

// batched (over alpha) 128-wide contraction  C[a][r][0..128) = sum_c Aop(a; r, c) * Bop(a; *, c)
int band_gemm(const mvb::GatherOperand& A, int64_t a_roff_batch, const mvb::GatherOperand& B, int64_t b_roff_batch,
              int64_t b_koff_batch, int rows, int batch, float* C, int cls, cudaStream_t st) {
  mvb::TcGemmParams g = mvb::tc_gemm_defaults();
  g.A = A; g.B = B; g.M = rows; g.N = 128; g.K = 128; g.C = C; g.ldc = 128; g.batch = batch;
  g.a_roff_batch = a_roff_batch; g.b_roff_batch = b_roff_batch; g.b_koff_batch = b_koff_batch;
  g.c_batch_stride = (int64_t)rows * 128;
  g_prof_class = cls;
  const int slot = prof_begin(2.0 * rows * 128.0 * 128.0 * batch, st);
  MVB_CUDA(mvb::tc_gemm_launch(g, 1, 128, st));
  prof_end(slot, st);
  g_prof_class = PC_OTHER;
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MVB_OK;
}

int ridge_solve_band(const mvb_design* d, const float* y, int F, float* G, const float* XtY, const float* mu, const float* ybar,
                     const float* alphas, int A, int fit_intercept, int per_target, float* coef, float* intercept,
                     float* alpha_out, int32_t* best, float* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  const int64_t AF = (int64_t)A * F;
  const int Tn = d->n_time;
  Arena ar(ws, ws_bytes);
  BandWs w;
  carve_band(ar, d, F, A, w);
  if (!ar.ok()) return fail(MVB_E_WORKSPACE, "ridge_solve(band): workspace %zu < %zu", ws_bytes, ar.off);
  const int64_t P = w.P, nb = w.nb;
  const int64_t rows_max = std::max<int64_t>(n, F);

  mvb::design_offsets_kernel<<<grid_for(n + p), 256, 0, st>>>(d->trial_idx, d->n_trials, Tn, d->trial_stride, d->feat_idx,
                                                            d->n_feat, d->n_lag, d->feat_stride, (int64_t)F * Tn,
                                                            w.row_off, w.col_off, w.yrow_off);
  MVB_LAUNCHED();
  const int64_t n_iota = std::max<int64_t>(P, AF) + 256;
  mvb::iota_offsets_kernel<<<grid_for(n_iota), 256, 0, st>>>(w.iota, n_iota, 1);                 MVB_LAUNCHED();
  const int64_t n_rowP = std::max<int64_t>(std::max<int64_t>(P, n), AF);
  mvb::iota_offsets_kernel<<<grid_for(n_rowP), 256, 0, st>>>(w.rowP, n_rowP, P);                 MVB_LAUNCHED();   // e * P
  const int64_t n_tab = std::max<int64_t>((int64_t)A * nb * 128, (int64_t)A * rows_max);
  mvb::iota_offsets_kernel<<<grid_for(n_tab), 256, 0, st>>>(w.tab128, n_tab, 128);               MVB_LAUNCHED();   // e * 128
  mvb::iota_offsets_kernel<<<grid_for(p), 256, 0, st>>>(w.iota_F, p, F);                         MVB_LAUNCHED();   // j * F
  const int64_t n_ip = std::max<int64_t>(p, AF);
  mvb::iota_offsets_kernel<<<grid_for(n_ip), 256, 0, st>>>(w.iota_p, n_ip, p);                   MVB_LAUNCHED();   // e * p
  pad_double_kernel<<<grid_for(P), 256, 0, st>>>(mu, p, P, w.mu64);                                MVB_LAUNCHED();
  pad_double_kernel<<<grid_for(F), 256, 0, st>>>(ybar, F, F, w.ybar64);                            MVB_LAUNCHED();

  // 1. Gp = Qt^T T Qt
  {
    cudaError_t err = cudaSuccess;
    mvb::bj::Hooks hooks{[](int, double flops, cudaStream_t s_) { g_prof_class = PC_EIG_REFRESH; return prof_begin(flops, s_); },
                         [](int slot, cudaStream_t s_) { prof_end(slot, s_); g_prof_class = PC_OTHER; }};
    const int nl = mvb::lb::reduce(G, (int)p, w.Qt, w.Tdiag, w.Tsub, w.lanczos_ws, st, &err, &hooks);
    if (nl < 0) return fail(MVB_E_CUDA, "band reduction failed: %s", cudaGetErrorString(err));
    g_launches.fetch_add(nl, std::memory_order_relaxed);
  }
  // 2. block Cholesky per alpha -> Minv, N, Pb, [Minv | -N].  Only 20 CTAs and independent of Z: runs on a side stream
  //    concurrently with the Z contraction and joins before the substitution sweep.
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  SideJoin side_join(st);
  {
    MVB_CUDA(cudaFuncSetAttribute(mvb::bl::block_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mvb::bl::CHOL_SMEM));
    cudaStream_t side = side_stream(st);
    if (!side) return fail(MVB_E_CUDA, "ridge_solve(band): no side stream for the current device");
    // one pair of events per host thread and device, reused: a wait binds to the record that precedes it in host order, so
    // reuse across calls and streams is safe, and nothing is created or destroyed while a caller is stream-capturing
    static thread_local cudaEvent_t tl_fork[64] = {}, tl_join[64] = {};
    int dev = 0;
    MVB_CUDA(cudaGetDevice(&dev));
    if (!tl_fork[dev]) {
      MVB_CUDA(cudaEventCreateWithFlags(&tl_fork[dev], cudaEventDisableTiming));
      MVB_CUDA(cudaEventCreateWithFlags(&tl_join[dev], cudaEventDisableTiming));
    }
    ev_fork = tl_fork[dev]; ev_join = tl_join[dev];
    MVB_CUDA(cudaEventRecord(ev_fork, st));
    MVB_CUDA(cudaStreamWaitEvent(side, ev_fork, 0));
    mvb::bl::block_chol_kernel<<<A, 1024, mvb::bl::CHOL_SMEM, side>>>(w.Tdiag, w.Tsub, (int)nb, alphas, w.Minv, w.Nmat, w.Pb, w.MN);
    MVB_LAUNCHED();
    MVB_CUDA(cudaEventRecord(ev_join, side));
    side_join.arm(ev_join);
  }
  // 3. Z = X_t Qt^T (implicit design rows x basis vectors) in column blocks [nb][n][128], centred: Z -= Qt mu
  const int qmode = dense_mode(P, (int)p, w.Qt, 4);
  mvb_operand opRows{d->xp, w.row_off, w.col_off, (int)n, 0};
  mvb_operand opQt{w.Qt, w.rowP, w.iota, (int)P, qmode};
  int rc;
  { ProfScope ps(PC_LEVERAGE_Z); rc = run_contract<float>(opRows, opQt, (int)n, (int)P, (int)p, w.Z, P, 0, w.contract_ws, w.contract_bytes, st, 128, n * 128, w.planes, w.planes_bytes); }
  if (rc) return rc;
  mvb::matvec_rows_kernel<float><<<cdiv(P * 32, 256), 256, 0, st>>>(w.Qt, P, P, w.mu64, w.cz);    MVB_LAUNCHED();
  sub_cols_blocked_kernel<<<grid_for(n * P), 256, 0, st>>>(w.Z, n, nb, w.cz);                        MVB_LAUNCHED();
  MVB_CUDA(side_join.join());
  // 4. forward substitution over the blocks for every (alpha, row):  u_j = [Minv_j | -N_j] [z_j ; u_{j-1}],  h = sum_j |u_j|^2.
  //    Default: one kernel in which every (row tile, alpha) CTA walks its whole chain with u resident on the SM
  //    (sweep_chain.cuh): in tensor memory as the MMA's A operand (MVB_SWEEP_CHAIN=2, default) or in shared memory (=1).
  //    MVB_SWEEP_CHAIN=0: one batched K = 256 contraction per block with |u|^2 in its epilogue.
  const char* chain_env = getenv("MVB_SWEEP_CHAIN");
  const int chain_mode = chain_env ? atoi(chain_env) : 2;
  const bool use_chain = chain_mode != 0;
  if (use_chain) {
    g_prof_class = PC_OTHER;
    const int slot = prof_begin(2.0 * n * 128.0 * 256.0 * A * nb, st);
    if (chain_mode == 2) MVB_CUDA(mvb::sc::sweep_chain_tmem_launch(w.Z, w.MN, n, (int)nb, A, w.h, st));
    else MVB_CUDA(mvb::sc::sweep_chain_launch(w.Z, w.MN, n, (int)nb, A, w.h, st));
    prof_end(slot, st);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    leverage_to_d_kernel<<<grid_for(n * A), 256, 0, st>>>(w.h, n, A, fit_intercept ? 1.f / (float)n : 0.f, w.dd, 1);   MVB_LAUNCHED();
  } else {
    MVB_CUDA(cudaMemsetAsync(w.h, 0, (size_t)mvb::TC_ROWSQ_PARTS * A * n * 4, st));
    sweep_koff_kernel<<<grid_for(nb * A * 256), 256, 0, st>>>(w.kofs, (int)nb, A, n, w.U0 - w.Z, w.U1 - w.Z);   MVB_LAUNCHED();
    mvb::iota_offsets_kernel<<<grid_for((int64_t)A * nb * 128), 256, 0, st>>>(w.tab256, (int64_t)A * nb * 128, 256);   MVB_LAUNCHED();
    for (int j = 0; j < nb; ++j) {
      mvb::TcGemmParams g = mvb::tc_gemm_defaults();
      g.A = mvb::GatherOperand{w.Z, w.tab128, w.kofs + (int64_t)j * A * 256, (int)n, 2};      // rows r -> r * 128
      g.B = mvb::GatherOperand{w.MN, w.tab256 + (int64_t)j * 128, w.iota, 128, 2};
      g.M = (int)n; g.N = 128; g.K = j > 0 ? 256 : 128; g.ldc = 128; g.batch = A;
      g.a_koff_batch = 256; g.b_roff_batch = nb * 128;
      g.C = (j & 1) ? w.U1 : w.U0;
      g.c_batch_stride = n * 128;
      g.rowsq = w.h;
      g_prof_class = PC_OTHER;
      const int slot = prof_begin(2.0 * n * 128.0 * g.K * A, st);
      MVB_CUDA(mvb::tc_gemm_launch(g, 1, 128, st));
      prof_end(slot, st);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    leverage_to_d_kernel<<<grid_for(n * A), 256, 0, st>>>(w.h, n, A, fit_intercept ? 1.f / (float)n : 0.f, w.dd, mvb::TC_ROWSQ_PARTS);   MVB_LAUNCHED();
  }
  // 5. beta for every alpha:  Ct[f] = Qt b_f ;  L u = Ct ;  L^T x = u ;  beta = Qt^T x
  {
    mvb_operand opB{XtY, w.iota, w.iota_F, F, 0};            // rows f (contiguous), contraction j (stride F)
    { ProfScope ps(PC_SMALL); rc = run_contract<float>(opB, opQt, F, (int)P, (int)p, w.Ct, P, 0, w.contract_ws, w.contract_bytes, st); }
    if (rc) return rc;
    // forward
    for (int j = 0; j < nb; ++j) {
      mvb::GatherOperand ac{w.Ct, w.rowP, w.iota + (int64_t)j * 128, F, 2};
      mvb::GatherOperand bm{w.Minv, w.tab128 + (int64_t)j * 128, w.iota, 128, 2};
      rc = band_gemm(ac, 0, bm, nb * 128, 0, F, A, w.C1, PC_OTHER, st);
      if (rc) return rc;
      if (j > 0) {
        mvb::GatherOperand au{w.Ub, w.rowP, w.iota + (int64_t)(j - 1) * 128, F, 2};
        mvb::GatherOperand bn{w.Nmat, w.tab128 + (int64_t)j * 128, w.iota, 128, 2};
        rc = band_gemm(au, F, bn, nb * 128, 0, F, A, w.C2, PC_OTHER, st);
        if (rc) return rc;
      }
      // Ub[a][f][j*128 ..] = C1 - C2 : rows (a,f) have ld P in Ub but C1 is [a][f][128]
      mvb::bl::recur_update_kernel<<<cdiv(AF * 32, 256), 256, 0, st>>>(w.C1, j > 0 ? w.C2 : nullptr, AF, w.Ub + (int64_t)j * 128, P, nullptr);
      MVB_LAUNCHED();
    }
    // backward: x_j = Minv_j^T u_j - Pb_j^T x_{j+1}
    for (int j = (int)nb - 1; j >= 0; --j) {
      mvb::GatherOperand au{w.Ub, w.rowP, w.iota + (int64_t)j * 128, F, 2};
      mvb::GatherOperand bmt{w.Minv, w.iota, w.tab128 + (int64_t)j * 128, 128, 0};     // transposed read: rows r, contraction c
      rc = band_gemm(au, F, bmt, 0, nb * 128, F, A, w.C1, PC_OTHER, st);
      if (rc) return rc;
      if (j < nb - 1) {
        mvb::GatherOperand ax{w.Xb, w.rowP, w.iota + (int64_t)(j + 1) * 128, F, 2};
        mvb::GatherOperand bpt{w.Pb, w.iota, w.tab128 + (int64_t)j * 128, 128, 0};
        rc = band_gemm(ax, F, bpt, 0, nb * 128, F, A, w.C2, PC_OTHER, st);
        if (rc) return rc;
      }
      mvb::bl::recur_update_kernel<<<cdiv(AF * 32, 256), 256, 0, st>>>(w.C1, j < nb - 1 ? w.C2 : nullptr, AF, w.Xb + (int64_t)j * 128, P, nullptr);
      MVB_LAUNCHED();
    }
  }
  // 6. fitted values for every (alpha, target) in the reduced basis:  X_c beta_a = Z_c (T + a I)^-1 Q b = Z_c Xb_a^T.  Z is already
  //    there (centred, column-blocked), so the coefficients beta = Q^T x are only formed for the selected alpha(s) afterwards
  //    (r01 built beta for the whole grid -- an (A F) x p x P product -- and contracted the implicit design with it).
  zblock_koff_kernel<<<grid_for(P), 256, 0, st>>>(w.zkoff, P, n);                                    MVB_LAUNCHED();
  {
    mvb_operand opZ{w.Z, w.tab128, w.zkoff, (int)n, 2};          // row i at i * 128 inside a column block, k -> block * n * 128 + k % 128
    mvb_operand opX{w.Xb, w.rowP, w.iota, (int)AF, 2};
    { ProfScope ps(PC_RESIDUAL); rc = run_contract<float>(opZ, opX, (int)n, (int)AF, (int)P, w.R, AF, 0, w.contract_ws, w.contract_bytes, st, 0, 0, w.planes, w.planes_bytes); }
    if (rc) return rc;
  }
  MVB_CUDA(cudaMemsetAsync(w.mb, 0, (size_t)AF * sizeof(double), st));        // Z is centred: no <beta, mu> term in the residuals
  {
    dim3 grid(cdiv(AF, 128), w.slabs);
    mvb::loo_loss_kernel<float><<<grid, 128, 0, st>>>(w.R, n, A, F, AF, y, w.yrow_off, (int64_t)Tn, w.ybar64, w.mb, w.dd, w.loss_partial);
    MVB_LAUNCHED();
    mvb::select_alpha_kernel<float><<<1, 256, 0, st>>>(w.loss_partial, w.slabs, A, F, 1.0 / (double)n, alphas, per_target, loss, best,
                                                       alpha_out, w.loss64);
    MVB_LAUNCHED();
  }
  // 7. coefficients of the chosen alpha(s): coef[f] = Q^T Xb[(best_f, f)], intercept = ybar - <coef, mu>
  {
    gather_rows_kernel<<<grid_for((int64_t)F * P), 256, 0, st>>>(w.Xb, F, P, best, per_target, w.Ssel);   MVB_LAUNCHED();
    mvb_operand opS{w.Ssel, w.rowP, w.iota, F, 2};
    const float* Qc = (const float*)((const char*)w.lanczos_ws + mvb::lb::make_plan((int)p).off_Qc);   // Q^T from the reduction
    mvb_operand opQc{Qc, w.rowP, w.iota, (int)p, 2};          // rows jj of Q^T, contraction k contiguous
    { ProfScope ps(PC_SMALL); rc = run_contract<float>(opS, opQc, F, (int)p, (int)P, coef, p, 0, w.contract_ws, w.contract_bytes, st); }
    if (rc) return rc;
    mvb::matvec_rows_kernel<float><<<cdiv((int64_t)F * 32, 256), 256, 0, st>>>(coef, F, p, w.mu64, w.mb);   MVB_LAUNCHED();
    intercept_kernel<<<cdiv(F, 256), 256, 0, st>>>(w.ybar64, w.mb, F, fit_intercept, intercept);            MVB_LAUNCHED();
  }
  return MVB_OK;
}

template <typename T>
int ridge_solve_impl(const mvb_design* d, const T* y, int F, T* G, const T* XtY, const T* mu, const T* ybar,
                     const T* alphas, int A, int fit_intercept, int per_target, T* coef, T* intercept, T* alpha_out,
                     int32_t* best, T* loss, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  const int64_t AF = (int64_t)A * F;
  const int Tn = d->n_time;
  if constexpr (sizeof(T) == 4) {
    if (use_band(p, MVB_F32))
      return ridge_solve_band(d, (const float*)y, F, (float*)G, (const float*)XtY, (const float*)mu, (const float*)ybar,
                              (const float*)alphas, A, fit_intercept, per_target, (float*)coef, (float*)intercept, (float*)alpha_out,
                              best, (float*)loss, ws, ws_bytes, st);
  }
  Arena ar(ws, ws_bytes);
  SolveWs w;
  carve_solve<T>(ar, d, F, A, w);
  if (!ar.ok()) return fail(MVB_E_WORKSPACE, "ridge_solve: workspace %zu < %zu", ws_bytes, ar.off);
  T *Vt = (T*)w.Vt, *lam = (T*)w.lam, *wt = (T*)w.w, *Vtb = (T*)w.Vtb, *theta = (T*)w.theta, *beta = (T*)w.beta,
    *ZR = (T*)w.ZR, *dd = (T*)w.d;

  mvb::design_offsets_kernel<<<grid_for(n + p), 256, 0, st>>>(d->trial_idx, d->n_trials, Tn, d->trial_stride, d->feat_idx,
                                                            d->n_feat, d->n_lag, d->feat_stride, (int64_t)F * Tn,
                                                            w.row_off, w.col_off, w.yrow_off);
  MVB_LAUNCHED();
  const int64_t niota = std::max<int64_t>(p, AF);
  mvb::iota_offsets_kernel<<<grid_for(niota), 256, 0, st>>>(w.iota_p, niota, p);   // i * p
  MVB_LAUNCHED();
  mvb::iota_offsets_kernel<<<grid_for(niota), 256, 0, st>>>(w.iota_1, niota, 1);   // i
  MVB_LAUNCHED();
  to_double_kernel<T><<<grid_for(p), 256, 0, st>>>(mu, p, w.mu64);
  MVB_LAUNCHED();
  to_double_kernel<T><<<grid_for(F), 256, 0, st>>>(ybar, F, w.ybar64);
  MVB_LAUNCHED();

  // 1. G = V diag(lam) V^T
  int rc = eigh_impl<T>(G, Vt, lam, 1, (int)p, w.eig_ws, w.eig_bytes, st);
  if (rc) return rc;
  // 2. ridge weights for the whole alpha grid
  mvb::ridge_weights_kernel<T><<<grid_for((int64_t)A * p), 256, 0, st>>>(lam, alphas, A, (int)p, wt);
  MVB_LAUNCHED();
  // 3. Vtb[k][f] = sum_j Vt[k][j] XtY[j][f]
  const int vmode = dense_mode(p, (int)p, Vt, sizeof(T));
  mvb_operand opVt{Vt, w.iota_p, w.iota_1, (int)p, vmode};                 // rows k, contraction j (contiguous)
  mvb::iota_offsets_kernel<<<grid_for(p), 256, 0, st>>>(w.iota_F, p, F);     // j * F
  MVB_LAUNCHED();
  mvb_operand opXtYt{XtY, w.iota_1, w.iota_F, F, 0};                       // rows f (contiguous), contraction j (stride F)
  { ProfScope ps(PC_SMALL); rc = run_contract<T>(opVt, opXtYt, (int)p, F, (int)p, Vtb, F, 0, w.contract_ws, w.contract_bytes, st); }
  if (rc) return rc;
  // 4. theta_t[(a,f)][k] = w[a][k] Vtb[k][f];  beta_all[(a,f)][j] = sum_k theta_t[(a,f)][k] Vt[k][j]
  mvb::theta_kernel<T><<<grid_for(AF * p), 256, 0, st>>>(wt, Vtb, A, F, (int)p, theta);
  MVB_LAUNCHED();
  mvb_operand opTheta{theta, w.iota_p, w.iota_1, (int)AF, dense_mode(p, (int)p, theta, sizeof(T))};
  mvb_operand opVcols{Vt, w.iota_1, w.iota_p, (int)p, 0};                  // rows j (contiguous), contraction k (stride p)
  { ProfScope ps(PC_SMALL); rc = run_contract<T>(opTheta, opVcols, (int)AF, (int)p, (int)p, beta, p, 0, w.contract_ws, w.contract_bytes, st); }
  if (rc) return rc;
  // 5. centring terms: cz[k] = <Vt[k], mu>,  mb[(a,f)] = <beta_all[(a,f)], mu>
  mvb::matvec_rows_kernel<T><<<cdiv(p * 32, 256), 256, 0, st>>>(Vt, p, p, w.mu64, w.cz);
  MVB_LAUNCHED();
  mvb::matvec_rows_kernel<T><<<cdiv(AF * 32, 256), 256, 0, st>>>(beta, AF, p, w.mu64, w.mb);
  MVB_LAUNCHED();
  // 6. Z = X_t V  (implicit design rows x eigenvectors), leverages -> d = 1 - h - c0
  mvb_operand opRows{d->xp, w.row_off, w.col_off, (int)n, 0};              // rows i (t contiguous), contraction j
  { ProfScope ps(PC_LEVERAGE_Z); rc = run_contract<T>(opRows, opVt, (int)n, (int)p, (int)p, ZR, p, 0, w.contract_ws, w.contract_bytes, st); }
  if (rc) return rc;
  mvb::leverage_kernel<T><<<cdiv(n, 64), 256, 0, st>>>(ZR, n, (int)p, p, w.cz, wt, A, fit_intercept ? 1.0 / (double)n : 0.0, dd);
  MVB_LAUNCHED();
  // 7. fitted values for every (alpha, target): R = X_t beta_all^T
  mvb_operand opBeta{beta, w.iota_p, w.iota_1, (int)AF, dense_mode(p, (int)p, beta, sizeof(T))};
  { ProfScope ps(PC_RESIDUAL); rc = run_contract<T>(opRows, opBeta, (int)n, (int)AF, (int)p, ZR, AF, 0, w.contract_ws, w.contract_bytes, st); }
  if (rc) return rc;
  // 8. LOO losses, alpha selection, coefficients
  {
    dim3 grid(cdiv(AF, 128), w.slabs);
    mvb::loo_loss_kernel<T><<<grid, 128, 0, st>>>(ZR, n, A, F, AF, y, w.yrow_off, (int64_t)Tn, w.ybar64, w.mb, dd, w.loss_partial);
    MVB_LAUNCHED();
    mvb::select_alpha_kernel<T><<<1, 256, 0, st>>>(w.loss_partial, w.slabs, A, F, 1.0 / (double)n, alphas, per_target, loss, best,
                                                   alpha_out, w.loss64);
    MVB_LAUNCHED();
    mvb::gather_coef_kernel<T><<<grid_for((int64_t)F * p), 256, 0, st>>>(beta, F, p, best, per_target, w.ybar64, w.mb, fit_intercept,
                                                                          coef, intercept);
    MVB_LAUNCHED();
  }
  return MVB_OK;
}

template <typename T>
int predict_impl(const mvb_design* d, const T* coef, const T* intercept, int F, T* yhat, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  constexpr int dt = sizeof(T) == 4 ? MVB_F32 : MVB_F64;
  Arena ar(ws, ws_bytes);
  int64_t* row_off = ar.take<int64_t>(n);
  int64_t* col_off = ar.take<int64_t>(p);
  int64_t* iota_p = ar.take<int64_t>(F);
  int64_t* iota_1 = ar.take<int64_t>(p);
  T* P = ar.take<T>(n * F);
  const size_t cb = plan_contract((int)n, F, (int)p, dt, 0).ws_elems * sizeof(T);
  void* cws = ar.take<char>(cb);
  if (!ar.ok()) return fail(MVB_E_WORKSPACE, "trf_predict: workspace %zu < %zu", ws_bytes, ar.off);
  if (!ws) return MVB_OK;
  mvb::design_offsets_kernel<<<grid_for(n + p), 256, 0, st>>>(d->trial_idx, d->n_trials, d->n_time, d->trial_stride, d->feat_idx,
                                                            d->n_feat, d->n_lag, d->feat_stride, 0, row_off, col_off, nullptr);
  MVB_LAUNCHED();
  mvb::iota_offsets_kernel<<<grid_for(F), 256, 0, st>>>(iota_p, F, p);
  MVB_LAUNCHED();
  mvb::iota_offsets_kernel<<<grid_for(p), 256, 0, st>>>(iota_1, p, 1);
  MVB_LAUNCHED();
  mvb_operand opRows{d->xp, row_off, col_off, (int)n, 0};
  mvb_operand opCoef{coef, iota_p, iota_1, F, dense_mode(p, (int)p, coef, sizeof(T))};
  int rc;
  { ProfScope ps(PC_PREDICT); rc = run_contract<T>(opRows, opCoef, (int)n, F, (int)p, P, F, 0, cws, cb, st); }
  if (rc) return rc;
  dim3 grid(cdiv(d->n_time, 32), cdiv(F, 32), d->n_trials), block(32, 8);
  mvb::predict_layout_kernel<T><<<grid, block, 0, st>>>(P, d->n_trials, F, d->n_time, F, intercept, yhat);
  MVB_LAUNCHED();
  return MVB_OK;
}


// ---- batched small dense LOO ridge (dense_batched.cuh) ------------------------------------------------------
struct DenseWs {
  float *Xc, *Yc, *mu, *ybar, *G, *XtY, *Li, *rowsq, *beta, *R, *d;
  float *dT, *eT, *tau, *V, *lfac, *dinv, *Q, *Z;  // tridiagonal route (Q, Z: explicit basis + tensor-core basis change)
  int64_t *roffQ, *koffQ, *roffZ;
  double* partial;
  int64_t *roffX, *roffXt, *koffN, *iota, *roffLi, *roffBeta;
  int Pp, tri_rows, tiles_per_alpha, slabs;
};

// 1: one Householder tridiagonalisation per slice (default); 0: one Cholesky factor + inverse + leverage product per (slice, alpha)
bool dense_route_tridiag(int p, int A) {
  static const bool on = [] { const char* e = getenv("MVB_DENSE_ROUTE"); return !e || strcmp(e, "chol") != 0; }();
  // the leverage kernel keeps the L D L^T factors of all alphas of a slice in shared memory next to a row tile: very long alpha grids
  // (2 A p floats beyond ~100 KB) take the Cholesky route, whose work is per alpha anyway
  const int Pp = (p + mvb::db::NB - 1) / mvb::db::NB * mvb::db::NB;
  return on && mvb::db::rows_apply_smem(Pp, p, A) <= (size_t)160 * 1024;
}

// tridiagonal route: 1 (default) = the reflectors are applied to the identity (Pp rows per slice) and Z = Xc Q comes from the batched
// tensor-core product; 0 = the reflectors are applied to every row tile of Xc (fp32 FMA work on n rows per slice)
bool dense_explicit_q() {
  static const bool on = [] { const char* e = getenv("MVB_DENSE_EXPLICIT_Q"); return !e || atoi(e) != 0; }();
  return on;
}

void carve_dense(Arena& ar, int B, int n, int p, int F, int A, DenseWs& w) {
  const int Pp = (p + mvb::db::NB - 1) / mvb::db::NB * mvb::db::NB;
  const int tri_rows = (Pp + mvb::db::TRI_TILE - 1) / mvb::db::TRI_TILE * mvb::db::TRI_TILE;
  w.Pp = Pp; w.tri_rows = tri_rows; w.tiles_per_alpha = tri_rows / mvb::db::TRI_TILE;
  w.slabs = 8;
  const int64_t AF = (int64_t)A * F;
  w.Xc = ar.take<float>((int64_t)B * n * Pp);
  w.Yc = ar.take<float>((int64_t)B * n * F);
  w.mu = ar.take<float>((int64_t)B * Pp);
  w.ybar = ar.take<float>((int64_t)B * F);
  w.G = ar.take<float>((int64_t)B * Pp * Pp);
  w.XtY = ar.take<float>((int64_t)B * Pp * F);
  const bool tri_route = dense_route_tridiag(p, A);
  w.Li = ar.take<float>(tri_route ? 0 : (int64_t)B * A * tri_rows * Pp);
  w.rowsq = ar.take<float>(tri_route ? 0 : (int64_t)A * w.tiles_per_alpha * mvb::TC_ROWSQ_PARTS * B * n);
  w.dT = ar.take<float>(tri_route ? (int64_t)B * Pp : 0);
  w.eT = ar.take<float>(tri_route ? (int64_t)B * Pp : 0);
  w.tau = ar.take<float>(tri_route ? (int64_t)B * Pp : 0);
  w.V = ar.take<float>(tri_route ? (int64_t)B * p * Pp : 0);
  w.lfac = ar.take<float>(tri_route ? (int64_t)B * A * Pp : 0);
  w.dinv = ar.take<float>(tri_route ? (int64_t)B * A * Pp : 0);
  const bool explicit_q = tri_route && dense_explicit_q();
  w.Q = ar.take<float>(explicit_q ? (int64_t)B * Pp * Pp : 0);
  w.Z = ar.take<float>(explicit_q ? (int64_t)B * n * Pp : 0);
  w.roffQ = ar.take<int64_t>(explicit_q ? (int64_t)B * Pp : 0);
  w.koffQ = ar.take<int64_t>(explicit_q ? Pp : 0);
  w.beta = ar.take<float>((int64_t)B * AF * Pp);
  w.R = ar.take<float>((int64_t)B * n * AF);
  w.d = ar.take<float>((int64_t)B * A * n);
  w.partial = ar.take<double>((int64_t)B * w.slabs * AF);
  w.roffX = ar.take<int64_t>((int64_t)B * n);
  w.roffXt = ar.take<int64_t>((int64_t)B * Pp);
  w.koffN = ar.take<int64_t>(n);
  w.iota = ar.take<int64_t>(std::max(n, Pp));
  w.roffLi = ar.take<int64_t>((int64_t)B * A * tri_rows);
  w.roffBeta = ar.take<int64_t>((int64_t)B * AF);
}

bool dense_batched_ok(int B, int n, int p, int F, int A) {
  return B >= 1 && p >= 1 && p <= mvb::db::MAX_COLS && F >= 1 && A >= 1 && n > p + 1 && (int64_t)A * F <= 4096 &&
         (int64_t)A * ((p + 127) / 128) * 128 <= 65535 * 128LL && n <= 65535;      // (n is a grid.z extent of the packing kernel)
}

int dense_batched_launch(mvb::TcGemmParams& g, int bn, int cls, double flops, cudaStream_t st) {
  g_prof_class = cls;
  const int slot = prof_begin(flops, st);
  cudaError_t e = mvb::tc_gemm_launch(g, 1, bn, st);
  prof_end(slot, st);
  g_prof_class = PC_OTHER;
  if (e != cudaSuccess) return fail(MVB_E_CUDA, "dense_ridge_batched: contraction launch failed: %s", cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MVB_OK;
}

int dense_ridge_batched_impl(const mvb_strided* X, const mvb_strided* Y, const int32_t* sample_idx, int n, int p, int F, int B,
                             const float* alphas, int A, int fit_intercept, int per_target, float* coef, float* intercept,
                             float* alpha_out, int32_t* best, float* loss, float* gram, int32_t* status, void* ws, size_t ws_bytes,
                             cudaStream_t st) {
  namespace db = mvb::db;
  Arena ar(ws, ws_bytes);
  DenseWs w;
  carve_dense(ar, B, n, p, F, A, w);
  if (!ar.ok()) return fail(MVB_E_WORKSPACE, "dense_ridge_batched: workspace %zu < %zu", ws_bytes, ar.off);
  const int Pp = w.Pp, AF = A * F;
  const db::View xv{(const float*)X->base, X->sample_stride, X->col_stride, X->batch_stride};
  const db::View yv{(const float*)Y->base, Y->sample_stride, Y->col_stride, Y->batch_stride};
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(db::chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)db::chol_inv_smem(db::MAX_COLS));
  });
  MVB_CUDA(attr_err);
  MVB_CUDA(cudaMemsetAsync(status, 0, (size_t)B * sizeof(int32_t), st));
  MVB_CUDA(cudaMemsetAsync(w.mu, 0, (size_t)B * Pp * sizeof(float), st));            // padding columns read as mean 0
  db::tables_kernel<<<grid_for((int64_t)B * (n + Pp + (int64_t)A * w.tri_rows + AF)), 256, 0, st>>>(B, n, Pp, AF, A, w.tri_rows, w.roffX, w.roffXt,
                                                                                                     w.koffN, w.iota, w.roffLi, w.roffBeta);
  MVB_LAUNCHED();
  // 1. centring before multiplying (ridgecv.py:59-94) + slice-major packing
  db::means_kernel<<<grid_for((int64_t)p * B), 256, 0, st>>>(xv, sample_idx, n, p, B, fit_intercept, w.mu, Pp);
  MVB_LAUNCHED();
  if ((int64_t)F * B < 65536) db::means_warp_kernel<<<cdiv((int64_t)F * B * 32, 256), 256, 0, st>>>(yv, sample_idx, n, F, B, fit_intercept, w.ybar, F);
  else db::means_kernel<<<grid_for((int64_t)F * B), 256, 0, st>>>(yv, sample_idx, n, F, B, fit_intercept, w.ybar, F);
  MVB_LAUNCHED();
  db::pack_kernel<<<dim3(cdiv(Pp, 32), cdiv(B, 32), n), dim3(32, 8), 0, st>>>(xv, sample_idx, n, p, B, w.mu, Pp, w.Xc, Pp);
  MVB_LAUNCHED();
  db::pack_kernel<<<dim3(cdiv(F, 32), cdiv(B, 32), n), dim3(32, 8), 0, st>>>(yv, sample_idx, n, F, B, w.ybar, F, w.Yc, F);
  MVB_LAUNCHED();
  // 2. G_b = Xc_b^T Xc_b (rows = columns of the slice, contraction over its samples), X^T y
  {
    mvb::TcGemmParams g = mvb::tc_gemm_defaults();
    g.A = {w.Xc, w.roffXt, w.koffN, Pp, 0};
    g.B = g.A;
    g.M = Pp; g.N = Pp; g.K = n; g.C = w.G; g.ldc = Pp; g.batch = B;
    g.a_roff_batch = Pp; g.b_roff_batch = Pp; g.c_batch_stride = (int64_t)Pp * Pp;
    int rc = dense_batched_launch(g, Pp > 128 ? 256 : 128, PC_GRAM, 2.0 * Pp * (double)Pp * n * B, st);
    if (rc) return rc;
  }
  db::xty_kernel<<<dim3(cdiv(Pp, 128), B, cdiv(F, 8)), 128, 0, st>>>(w.Xc, w.Yc, n, Pp, F, w.XtY);
  MVB_LAUNCHED();
  if (gram) {
    db::copy_gram_kernel<<<grid_for((int64_t)B * p * p), 256, 0, st>>>(w.G, B, p, Pp, gram);
    MVB_LAUNCHED();
  }
  if (dense_route_tridiag(p, A)) {
    // 3'. one orthogonal reduction per slice: G = Q T Q^T;  Q^T X^T y;  L D L^T of T + alpha I and the reduced-basis solutions;
    //     Z = Xc Q in place with the leverages of every alpha from the tridiagonal factors
    static std::once_flag tri_once;
    static cudaError_t tri_err = cudaSuccess;
    std::call_once(tri_once, [] {
      tri_err = cudaFuncSetAttribute(db::tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)db::tridiag_smem(db::MAX_COLS));
      if (tri_err == cudaSuccess)
        tri_err = cudaFuncSetAttribute(db::rows_apply_leverage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    });
    MVB_CUDA(tri_err);
    const size_t ra_smem = db::rows_apply_smem(Pp, p, A);
    db::tridiag_kernel<<<B, db::TD_THREADS, db::tridiag_smem(p), st>>>(w.G, p, Pp, w.dT, w.eT, Pp, w.V, Pp, w.tau);
    MVB_LAUNCHED();
    db::reflectors_small_kernel<<<dim3(cdiv(F, 8), B), 256, 0, st>>>(w.V, Pp, w.tau, Pp, p, w.XtY, (int64_t)Pp * F, F, 1, F, 0);
    MVB_LAUNCHED();
    db::tri_ldl_solve_kernel<<<cdiv((int64_t)B * A, 128), 128, 0, st>>>(w.dT, w.eT, Pp, p, Pp, alphas, A, B, w.XtY, F, 1e-4f, w.lfac, w.dinv, w.beta, status);
    MVB_LAUNCHED();
    const float c0 = fit_intercept ? 1.f / (float)n : 0.f;
    float* Zrows = w.Xc;
    if (dense_explicit_q()) {
      // Q = I H_0 H_1 ... (the reflectors applied to Pp identity rows per slice), then Z = Xc Q on the tensor cores, then the leverages
      db::identity_batched_kernel<<<grid_for((int64_t)B * Pp * Pp), 256, 0, st>>>(w.Q, B, Pp);
      MVB_LAUNCHED();
      db::rows_apply_leverage_kernel<<<dim3(cdiv(Pp, db::RA_ROWS), B), 256, ra_smem, st>>>(w.Q, Pp, p, Pp, w.V, Pp, w.tau, Pp, w.lfac, w.dinv, 0, 0.f,
                                                                                            nullptr, 1);
      MVB_LAUNCHED();
      db::dense_q_tables_kernel<<<grid_for((int64_t)B * Pp), 256, 0, st>>>(B, Pp, w.roffQ, w.koffQ);
      MVB_LAUNCHED();
      mvb::TcGemmParams g = mvb::tc_gemm_defaults();
      g.A = {w.Xc, w.roffX, w.iota, n, 2};
      g.B = {w.Q, w.roffQ, w.koffQ, Pp, 0};                // B(k', j) = Q[j][k']: rows k' contiguous, contraction j strides Pp
      g.M = n; g.N = Pp; g.K = Pp; g.C = w.Z; g.ldc = Pp; g.batch = B;
      g.a_roff_batch = n; g.b_roff_batch = Pp; g.c_batch_stride = (int64_t)n * Pp;
      int rc = dense_batched_launch(g, Pp > 128 ? 256 : 128, PC_LEVERAGE_Z, 2.0 * n * (double)Pp * Pp * B, st);
      if (rc) return rc;
      Zrows = w.Z;
      db::rows_apply_leverage_kernel<<<dim3(cdiv(n, db::RA_ROWS), B), 256, ra_smem, st>>>(w.Z, n, p, Pp, w.V, Pp, w.tau, Pp, w.lfac, w.dinv, A, c0, w.d, 0);
      MVB_LAUNCHED();
    } else {
      db::rows_apply_leverage_kernel<<<dim3(cdiv(n, db::RA_ROWS), B), 256, ra_smem, st>>>(w.Xc, n, p, Pp, w.V, Pp, w.tau, Pp, w.lfac, w.dinv, A, c0, w.d, 1);
      MVB_LAUNCHED();
    }
    {
      mvb::TcGemmParams g = mvb::tc_gemm_defaults();
      g.A = {Zrows, w.roffX, w.iota, n, 2};                // Z (in place of Xc, or the product's output: same row layout)
      g.B = {w.beta, w.roffBeta, w.iota, AF, 2};
      g.M = n; g.N = AF; g.K = Pp; g.C = w.R; g.ldc = AF; g.batch = B;
      g.a_roff_batch = n; g.b_roff_batch = AF; g.c_batch_stride = (int64_t)n * AF;
      int rc = dense_batched_launch(g, AF > 128 ? 256 : 128, PC_RESIDUAL, 2.0 * n * (double)AF * Pp * B, st);
      if (rc) return rc;
    }
    db::loss_kernel<<<dim3(B, w.slabs), 128, 0, st>>>(w.R, w.Yc, w.d, n, A, F, w.slabs, w.partial);
    MVB_LAUNCHED();
    db::select_kernel<<<B, 256, (size_t)(AF + F) * 4, st>>>(w.partial, w.slabs, n, A, F, p, Pp, alphas, per_target, fit_intercept, w.beta, nullptr,
                                                           w.ybar, loss, best, alpha_out, coef, intercept);
    MVB_LAUNCHED();
    db::reflectors_small_kernel<<<dim3(cdiv(F, 8), B), 256, 0, st>>>(w.V, Pp, w.tau, Pp, p, coef, (int64_t)F * p, 1, p, F, 1);      // coef = Q x
    MVB_LAUNCHED();
    db::intercept_batched_kernel<<<cdiv((int64_t)B * F * 32, 256), 256, 0, st>>>(coef, w.mu, w.ybar, B, F, p, Pp, fit_intercept, intercept);
    MVB_LAUNCHED();
    return MVB_OK;
  }
  // 3. per (slice, alpha): Cholesky factor and its inverse
  db::chol_inv_kernel<<<B * A, db::CI_THREADS, db::chol_inv_smem(Pp), st>>>(w.G, p, Pp, alphas, A, 1e-4f, w.Li, w.tri_rows, status);
  MVB_LAUNCHED();
  // 4. leverages: row squares of Xc_b [L_a^-1]^T for the stack of all alphas (nothing but the squares is stored)
  {
    mvb::TcGemmParams g = mvb::tc_gemm_defaults();
    g.A = {w.Xc, w.roffX, w.iota, n, 2};
    g.B = {w.Li, w.roffLi, w.iota, A * w.tri_rows, 2};
    g.M = n; g.N = A * w.tri_rows; g.K = Pp; g.C = nullptr; g.ldc = 0; g.batch = B;
    g.a_roff_batch = n; g.b_roff_batch = (int64_t)A * w.tri_rows;
    g.rowsq = w.rowsq; g.rowsq_assign = 1; g.b_tri_period = w.tiles_per_alpha;      // every (tile, part, slice, row) entry is written once
    // a CTA walks the N tiles of `walk` alphas for its 128 rows (barriers / tensor memory set up once, the A rows stay in L1 / L2)
    static const int walk = [] { const char* e = getenv("MVB_DENSE_WALK"); return e ? atoi(e) : 4; }();
    g.n_tiles_per_cta = walk > 1 ? walk * w.tiles_per_alpha : 0;
    // flops counted for the triangular half actually contracted
    double kk = 0;
    for (int t = 0; t < w.tiles_per_alpha; ++t) kk += std::min(Pp, (t + 1) * db::TRI_TILE);
    int rc = dense_batched_launch(g, db::TRI_TILE, PC_LEVERAGE_Z, 2.0 * n * (double)db::TRI_TILE * kk * A * B, st);
    if (rc) return rc;
  }
  db::leverage_d_kernel<<<grid_for((int64_t)B * A * n), 256, 0, st>>>(w.rowsq, B, A, n, w.tiles_per_alpha, fit_intercept ? 1.f / (float)n : 0.f, w.d);
  MVB_LAUNCHED();
  // 5. coefficients of every alpha, fitted values, LOO losses, selection
  db::beta_kernel<<<B * A, 256, 0, st>>>(w.Li, w.tri_rows, Pp, w.XtY, F, A, w.beta);
  MVB_LAUNCHED();
  {
    mvb::TcGemmParams g = mvb::tc_gemm_defaults();
    g.A = {w.Xc, w.roffX, w.iota, n, 2};
    g.B = {w.beta, w.roffBeta, w.iota, AF, 2};
    g.M = n; g.N = AF; g.K = Pp; g.C = w.R; g.ldc = AF; g.batch = B;
    g.a_roff_batch = n; g.b_roff_batch = AF; g.c_batch_stride = (int64_t)n * AF;
    int rc = dense_batched_launch(g, AF > 128 ? 256 : 128, PC_RESIDUAL, 2.0 * n * (double)AF * Pp * B, st);
    if (rc) return rc;
  }
  db::loss_kernel<<<dim3(B, w.slabs), 128, 0, st>>>(w.R, w.Yc, w.d, n, A, F, w.slabs, w.partial);
  MVB_LAUNCHED();
  db::select_kernel<<<B, 256, (size_t)(AF + F) * 4, st>>>(w.partial, w.slabs, n, A, F, p, Pp, alphas, per_target, fit_intercept, w.beta, w.mu,
                                                         w.ybar, loss, best, alpha_out, coef, intercept);
  MVB_LAUNCHED();
  return MVB_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int mvb_version(void) { return 100; }
const char* mvb_last_error(void) { return g_err; }
int64_t mvb_launch_count(void) { return g_launches.load(); }

void mvb_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int c = 0; c < PC_COUNT; ++c) { g_prof.n_timed[c] = 0; g_prof.n_all[c] = 0; g_prof.flops_timed[c] = 0; g_prof.flops_all[c] = 0; }
  g_prof.on = true;
}
void mvb_profile_disable(void) { g_prof.on = false; }
int mvb_profile_set(int on) { const int was = g_prof.on ? 1 : 0; g_prof.on = on != 0; return was; }
void mvb_launch_count_add(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int mvb_profile_read(double* ms_timed, double* flops_timed, double* flops_all, int64_t* n_timed, int64_t* n_all, int max_classes) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  const int nc = max_classes < PC_COUNT ? max_classes : PC_COUNT;
  for (int c = 0; c < nc; ++c) {
    double ms = 0;
    for (int i = 0; i < g_prof.n_timed[c]; ++i) {
      float t = 0.f;
      if (cudaEventSynchronize(g_prof.ev[c][i][1]) == cudaSuccess && cudaEventElapsedTime(&t, g_prof.ev[c][i][0], g_prof.ev[c][i][1]) == cudaSuccess) ms += t;
    }
    ms_timed[c] = ms; flops_timed[c] = g_prof.flops_timed[c]; flops_all[c] = g_prof.flops_all[c];
    n_timed[c] = g_prof.n_timed[c]; n_all[c] = g_prof.n_all[c];
  }
  return nc;
}
const char* mvb_profile_name(int cls) { return (cls >= 0 && cls < PC_COUNT) ? kProfNames[cls] : ""; }

int mvb_scaler_fit(const void* X, int64_t R, int64_t K, int dtype, void* mean, void* var, void* scale, void* stream) {
  if (!X || !mean || !var || !scale || R < 1 || K < 1) return fail(MVB_E_ARG, "scaler_fit: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MVB_F32) mvb::scaler_fit_kernel<float><<<cdiv(K, 128), 128, 0, st>>>((const float*)X, R, K, (float*)mean, (float*)var, (float*)scale);
  else if (dtype == MVB_F64) mvb::scaler_fit_kernel<double><<<cdiv(K, 128), 128, 0, st>>>((const double*)X, R, K, (double*)mean, (double*)var, (double*)scale);
  else return fail(MVB_E_UNSUPPORTED, "scaler_fit: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_scaler_apply(const void* X, int64_t R, int64_t K, int dtype, const void* mean, const void* scale, int with_mean,
                     int with_std, int inverse, void* out, void* stream) {
  if (!X || !out || R < 1 || K < 1 || (with_mean && !mean) || (with_std && !scale)) return fail(MVB_E_ARG, "scaler_apply: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = R * K;
  if (dtype == MVB_F32) mvb::scaler_apply_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)X, total, K, (const float*)mean, (const float*)scale, with_mean, with_std, inverse, (float*)out);
  else if (dtype == MVB_F64) mvb::scaler_apply_kernel<double><<<grid_for(total), 256, 0, st>>>((const double*)X, total, K, (const double*)mean, (const double*)scale, with_mean, with_std, inverse, (double*)out);
  else return fail(MVB_E_UNSUPPORTED, "scaler_apply: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_trf_pad(const void* X, int N, int C, int T, int left, int right, int dtype, void* Xp, void* stream) {
  if (!X || !Xp || N < 1 || C < 1 || T < 1 || left < 0 || right < 0) return fail(MVB_E_ARG, "trf_pad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = (int64_t)N * C;
  const int Tp = T + left + right;
  if (dtype == MVB_F32) mvb::pad_kernel<float><<<grid_for(rows * Tp), 256, 0, st>>>((const float*)X, rows, T, left, Tp, (float*)Xp);
  else if (dtype == MVB_F64) mvb::pad_kernel<double><<<grid_for(rows * Tp), 256, 0, st>>>((const double*)X, rows, T, left, Tp, (double*)Xp);
  else return fail(MVB_E_UNSUPPORTED, "trf_pad: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_trf_shift(const void* xp, int N, int C, int Tp, const int32_t* trial_idx, int n_trials, int dtype, void* shift, void* out,
                  void* stream) {
  if (!xp || !shift || !out || N < 1 || C < 1 || Tp < 1 || n_trials < 1) return fail(MVB_E_ARG, "trf_shift: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)N * C * Tp;
  if (dtype == MVB_F32) {
    mvb::feature_pivot_kernel<float><<<C, 256, 0, st>>>((const float*)xp, trial_idx, n_trials, (int64_t)C * Tp, Tp, Tp, (float*)shift);
    MVB_LAUNCHED();
    mvb::feature_shift_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)xp, total, C, Tp, (const float*)shift, (float*)out);
  } else if (dtype == MVB_F64) {
    mvb::feature_pivot_kernel<double><<<C, 256, 0, st>>>((const double*)xp, trial_idx, n_trials, (int64_t)C * Tp, Tp, Tp, (double*)shift);
    MVB_LAUNCHED();
    mvb::feature_shift_kernel<double><<<grid_for(total), 256, 0, st>>>((const double*)xp, total, C, Tp, (const double*)shift, (double*)out);
  } else return fail(MVB_E_UNSUPPORTED, "trf_shift: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_transpose(const void* X, int64_t R, int64_t C, int dtype, void* out, void* stream) {
  if (!X || !out || R < 1 || C < 1) return fail(MVB_E_ARG, "transpose: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(C, 32), cdiv(R, 32)), block(32, 8);
  if (dtype == MVB_F32) mvb::transpose_kernel<float><<<grid, block, 0, st>>>((const float*)X, R, C, (float*)out);
  else if (dtype == MVB_F64) mvb::transpose_kernel<double><<<grid, block, 0, st>>>((const double*)X, R, C, (double*)out);
  else return fail(MVB_E_UNSUPPORTED, "transpose: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

size_t mvb_contract_workspace_bytes(int M, int N, int K, int dtype) {
  const size_t a = plan_contract(M, N, K, dtype, 0).ws_elems, b = plan_contract(M, N, K, dtype, 1).ws_elems;
  return std::max(a, b) * (dtype == MVB_F32 ? 4 : 8) + 256;
}

int mvb_contract(const mvb_operand* A, const mvb_operand* B, int M, int N, int K, int dtype, void* C, int64_t ldc,
                 int symmetric, void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !C || !A->base || !B->base || !A->roff || !A->koff || !B->roff || !B->koff || M < 0 || N < 0 || K < 0 || ldc < N)
    return fail(MVB_E_ARG, "contract: bad argument");
  if (A->rows < M || B->rows < N) return fail(MVB_E_ARG, "contract: operand rows smaller than M/N");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MVB_F32) return run_contract<float>(*A, *B, M, N, K, (float*)C, ldc, symmetric, ws, ws_bytes, st);
  if (dtype == MVB_F64) return run_contract<double>(*A, *B, M, N, K, (double*)C, ldc, symmetric, ws, ws_bytes, st);
  return fail(MVB_E_UNSUPPORTED, "contract: dtype %d", dtype);
}

size_t mvb_trf_stats_workspace_bytes(const mvb_design* d, int n_targets) {
  if (!design_ok(d) || n_targets < 1) return 0;
  Arena ar(nullptr, 0);
  StatsWs w;
  if (d->dtype == MVB_F32) carve_stats<float>(ar, d, n_targets, w); else carve_stats<double>(ar, d, n_targets, w);
  return ar.off + 256;
}

int mvb_trf_stats(const mvb_design* d, const void* y, int n_targets, int fit_intercept, void* G, void* XtY, void* mu,
                  void* ybar, void* ws, size_t ws_bytes, void* stream) {
  if (!design_ok(d) || !y || n_targets < 1 || !G || !XtY || !mu || !ybar || !ws) return fail(MVB_E_ARG, "trf_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == MVB_F32) return trf_stats_impl<float>(d, (const float*)y, n_targets, fit_intercept, (float*)G, (float*)XtY, (float*)mu, (float*)ybar, ws, ws_bytes, st);
  return trf_stats_impl<double>(d, (const double*)y, n_targets, fit_intercept, (double*)G, (double*)XtY, (double*)mu, (double*)ybar, ws, ws_bytes, st);
}

int mvb_gather_stats(const void* G, const void* XtY, const void* mu, int64_t p_full, int n_targets, const int32_t* feat_idx, int n_feat,
                     int n_lag, int dtype, void* G_sub, void* XtY_sub, void* mu_sub, void* stream) {
  if (!G || !XtY || !mu || !feat_idx || !G_sub || !XtY_sub || !mu_sub || p_full < 1 || n_targets < 1 || n_feat < 1 || n_lag < 1)
    return fail(MVB_E_ARG, "gather_stats: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ps = (int64_t)n_feat * n_lag;
  const int grid = grid_for(ps * ps + ps * n_targets + ps);
  if (dtype == MVB_F32) mvb::gather_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)G, (const float*)XtY, (const float*)mu, p_full, n_targets, feat_idx, n_feat, n_lag, (float*)G_sub, (float*)XtY_sub, (float*)mu_sub);
  else if (dtype == MVB_F64) mvb::gather_stats_kernel<double><<<grid, 256, 0, st>>>((const double*)G, (const double*)XtY, (const double*)mu, p_full, n_targets, feat_idx, n_feat, n_lag, (double*)G_sub, (double*)XtY_sub, (double*)mu_sub);
  else return fail(MVB_E_UNSUPPORTED, "gather_stats: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

size_t mvb_eigh_workspace_bytes(int batch, int p, int dtype) { return eigh_ws_bytes(batch, p, dtype); }

int mvb_eigh(void* A, void* Vt, void* lam, int batch, int p, int dtype, void* ws, size_t ws_bytes, void* stream) {
  if (!A || !Vt || !lam || batch < 1 || p < 1 || !ws) return fail(MVB_E_ARG, "eigh: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MVB_F32) return eigh_impl<float>((float*)A, (float*)Vt, (float*)lam, batch, p, ws, ws_bytes, st);
  if (dtype == MVB_F64) return eigh_impl<double>((double*)A, (double*)Vt, (double*)lam, batch, p, ws, ws_bytes, st);
  return fail(MVB_E_UNSUPPORTED, "eigh: dtype %d", dtype);
}

int mvb_band_padded(int p) { return p < 1 ? 0 : mvb::lb::make_plan(p).P; }
size_t mvb_band_reduce_workspace_bytes(int p) { return p < 1 ? 0 : mvb::lb::make_plan(p).bytes + 256; }

int mvb_band_reduce(const void* G, int p, void* Qt, void* Tdiag, void* Tsub, void* ws, size_t ws_bytes, void* stream) {
  if (!G || p < 1 || !Qt || !Tdiag || !Tsub || !ws) return fail(MVB_E_ARG, "band_reduce: bad argument");
  if (ws_bytes < mvb::lb::make_plan(p).bytes) return fail(MVB_E_WORKSPACE, "band_reduce: workspace %zu < %zu", ws_bytes, mvb::lb::make_plan(p).bytes);
  cudaError_t err = cudaSuccess;
  const int nl = mvb::lb::reduce((const float*)G, p, (float*)Qt, (float*)Tdiag, (float*)Tsub, ws, (cudaStream_t)stream, &err, nullptr);
  if (nl < 0) return fail(MVB_E_CUDA, "band reduction failed: %s", cudaGetErrorString(err));
  g_launches.fetch_add(nl, std::memory_order_relaxed);
  return MVB_OK;
}

int mvb_band_reduce_counters(const void* ws, int p, unsigned int* host_out) {
  if (!ws || p < 1 || !host_out) return fail(MVB_E_ARG, "band_reduce_counters: bad argument");
  const mvb::lb::Plan pl = mvb::lb::make_plan(p);
  mvb::bj::Ctl ctl;
  MVB_CUDA(cudaMemcpy(&ctl, (const char*)ws + pl.off_ctl, sizeof(ctl), cudaMemcpyDeviceToHost));   // synchronising: diagnostics only
  host_out[0] = ctl.rot[2]; host_out[1] = ctl.rot[3];
  return MVB_OK;
}

size_t mvb_ridge_solve_workspace_bytes(const mvb_design* d, int n_targets, int n_alphas) {
  if (!design_ok(d) || n_targets < 1 || n_alphas < 1) return 0;
  Arena ar(nullptr, 0);
  const int64_t p = (int64_t)d->n_feat * d->n_lag;
  if (use_band(p, d->dtype)) { BandWs bw; carve_band(ar, d, n_targets, n_alphas, bw); return ar.off + 256; }
  SolveWs w;
  if (d->dtype == MVB_F32) carve_solve<float>(ar, d, n_targets, n_alphas, w); else carve_solve<double>(ar, d, n_targets, n_alphas, w);
  return ar.off + 256;
}

int mvb_ridge_solve(const mvb_design* d, const void* y, int n_targets, void* G, const void* XtY, const void* mu,
                    const void* ybar, const void* alphas, int n_alphas, int fit_intercept, int alpha_per_target,
                    void* coef, void* intercept, void* alpha_out, int32_t* best, void* loss, void* ws, size_t ws_bytes,
                    void* stream) {
  if (!design_ok(d) || !y || n_targets < 1 || !G || !XtY || !mu || !ybar || !alphas || n_alphas < 1 || !coef || !intercept ||
      !alpha_out || !best || !ws)
    return fail(MVB_E_ARG, "ridge_solve: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == MVB_F32)
    return ridge_solve_impl<float>(d, (const float*)y, n_targets, (float*)G, (const float*)XtY, (const float*)mu, (const float*)ybar,
                                   (const float*)alphas, n_alphas, fit_intercept, alpha_per_target, (float*)coef, (float*)intercept,
                                   (float*)alpha_out, best, (float*)loss, ws, ws_bytes, st);
  return ridge_solve_impl<double>(d, (const double*)y, n_targets, (double*)G, (const double*)XtY, (const double*)mu,
                                  (const double*)ybar, (const double*)alphas, n_alphas, fit_intercept, alpha_per_target,
                                  (double*)coef, (double*)intercept, (double*)alpha_out, best, (double*)loss, ws, ws_bytes, st);
}

size_t mvb_trf_predict_workspace_bytes(const mvb_design* d, int n_targets) {
  if (!design_ok(d) || n_targets < 1) return 0;
  const int64_t n = (int64_t)d->n_trials * d->n_time, p = (int64_t)d->n_feat * d->n_lag;
  const size_t es = d->dtype == MVB_F32 ? 4 : 8;
  size_t b = align_up(n * 8) + align_up(p * 8) + align_up((size_t)n_targets * 8) + align_up(p * 8) + align_up(n * n_targets * es) +
             align_up(plan_contract((int)n, n_targets, (int)p, d->dtype, 0).ws_elems * es);
  return b + 1024;
}

int mvb_trf_predict(const mvb_design* d, const void* coef, const void* intercept, int n_targets, void* yhat, void* ws,
                    size_t ws_bytes, void* stream) {
  if (!design_ok(d) || !coef || n_targets < 1 || !yhat || !ws) return fail(MVB_E_ARG, "trf_predict: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->dtype == MVB_F32) return predict_impl<float>(d, (const float*)coef, (const float*)intercept, n_targets, (float*)yhat, ws, ws_bytes, st);
  return predict_impl<double>(d, (const double*)coef, (const double*)intercept, n_targets, (double*)yhat, ws, ws_bytes, st);
}

int mvb_dense_ridge_batched_max_cols(void) { return mvb::db::MAX_COLS; }

// developer entry (not in include/mvb.h; tools/skinny_time.py): one skinny Lanczos-shaped product C (128 x N) = A (128 x K) B (N x K)^T on the
// TMEM-A kernel with B as pre-split planes.  planes must hold tc_planes_bytes(N, K, raw); part: [splits][128][N] floats; tables: rowK[max(N,128)]
// = i * K, iota[K].  mode bit 0: raw single planes, bit 1: stream-once cache hints.
int mvb_dev_skinny(const float* A, const float* B, int N, int K, int splits, int mode, void* planes, float* part, const int64_t* rowK,
                   const int64_t* iota, int presplit, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  mvb::GatherOperand a{A, rowK, iota, 128, 2}, b{B, rowK, iota, N, 2};
  const bool raw = mode & 1;
  if (presplit) { MVB_CUDA(mvb::tc_presplit_launch(b, K, (uint8_t*)planes, st, 0, -1, 0, -1, raw)); return MVB_OK; }
  mvb::TcGemmParams g = mvb::tc_gemm_defaults();
  g.A = a; g.B = b; g.M = 128; g.N = N; g.K = K; g.C = part; g.ldc = N; g.c_split_stride = (int64_t)128 * N;
  g.b_planes = (const uint8_t*)planes; g.b_planes_nkb = (K + mvb::tc::BK - 1) / mvb::tc::BK; g.b_planes_raw = raw ? 1 : 0;
  g.b_stream_once = (mode & 2) ? 1 : 0;
  MVB_CUDA(mvb::tc_gemm_launch(g, splits, 256, st));
  return MVB_OK;
}

// developer entry (not in include/mvb.h; tools/tri_phases.py): per-phase cycles of block (0, 0) of the last rows_apply_leverage_kernel launch
int mvb_dev_rows_apply_phases(long long* host_out) {
  MVB_CUDA(cudaDeviceSynchronize());
  MVB_CUDA(cudaMemcpyFromSymbol(host_out, mvb::db::g_ra_clk, 4 * sizeof(long long)));
  MVB_CUDA(cudaMemcpyFromSymbol(host_out + 4, mvb::db::g_td_clk, 8 * sizeof(long long)));      // + the tridiagonalisation's phases
  return MVB_OK;
}

// developer entry (not in include/mvb.h; tools/chol_inv_phases.py): the factorisation kernel alone with per-phase cycle counters
int mvb_dev_chol_inv(const float* G, int p, int batch, const float* alphas, int n_alphas, float* Li, int* status, long long* phase_clk,
                     void* stream) {
  namespace db = mvb::db;
  const int Pp = (p + db::NB - 1) / db::NB * db::NB, tri_rows = (Pp + db::TRI_TILE - 1) / db::TRI_TILE * db::TRI_TILE;
  if (p < 1 || p > db::MAX_COLS) return fail(MVB_E_ARG, "dev_chol_inv: p");
  MVB_CUDA(cudaFuncSetAttribute(db::chol_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)db::chol_inv_smem(db::MAX_COLS)));
  db::chol_inv_kernel<<<batch * n_alphas, db::CI_THREADS, db::chol_inv_smem(Pp), (cudaStream_t)stream>>>(G, p, Pp, alphas, n_alphas, 1e-4f, Li, tri_rows,
                                                                                                     status, phase_clk);
  MVB_LAUNCHED();
  return MVB_OK;
}

size_t mvb_dense_ridge_batched_workspace_bytes(int batch, int n_samples, int n_cols, int n_targets, int n_alphas) {
  if (!dense_batched_ok(batch, n_samples, n_cols, n_targets, n_alphas) || batch > 65535) return 0;
  Arena ar(nullptr, 0);
  DenseWs w;
  carve_dense(ar, batch, n_samples, n_cols, n_targets, n_alphas, w);
  return ar.off + 256;
}

int mvb_dense_ridge_batched(const mvb_strided* X, const mvb_strided* Y, const int32_t* sample_idx, int n_samples, int n_cols,
                            int n_targets, int batch, const void* alphas, int n_alphas, int fit_intercept, int alpha_per_target,
                            int dtype, void* coef, void* intercept, void* alpha_out, int32_t* best, void* loss, void* gram,
                            int32_t* status, void* ws, size_t ws_bytes, void* stream) {
  if (!X || !Y || !X->base || !Y->base || !alphas || !coef || !intercept || !alpha_out || !best || !status || !ws)
    return fail(MVB_E_ARG, "dense_ridge_batched: bad argument");
  if (dtype != MVB_F32) return fail(MVB_E_UNSUPPORTED, "dense_ridge_batched: float32 only (float64 problems take mvb_ridge_solve per problem)");
  if (!dense_batched_ok(batch, n_samples, n_cols, n_targets, n_alphas) || batch > 65535)
    return fail(MVB_E_UNSUPPORTED, "dense_ridge_batched: needs 1 <= n_cols <= %d, n_samples > n_cols + 1, batch <= 65535 (got n = %d, p = %d, batch = %d)",
                mvb::db::MAX_COLS, n_samples, n_cols, batch);
  return dense_ridge_batched_impl(X, Y, sample_idx, n_samples, n_cols, n_targets, batch, (const float*)alphas, n_alphas, fit_intercept,
                                  alpha_per_target, (float*)coef, (float*)intercept, (float*)alpha_out, best, (float*)loss, (float*)gram,
                                  status, ws, ws_bytes, (cudaStream_t)stream);
}

int mvb_dense_predict_batched(const mvb_strided* X, int n_samples, int n_cols, int batch, const void* coef, const void* intercept,
                              int n_targets, int dtype, void* yhat, int64_t out_sample_stride, int64_t out_target_stride,
                              int64_t out_batch_stride, void* stream) {
  if (!X || !X->base || !coef || !yhat || n_samples < 1 || n_cols < 1 || batch < 1 || n_targets < 1)
    return fail(MVB_E_ARG, "dense_predict_batched: bad argument");
  if (dtype != MVB_F32) return fail(MVB_E_UNSUPPORTED, "dense_predict_batched: float32 only");
  cudaStream_t st = (cudaStream_t)stream;
  const mvb::db::View xv{(const float*)X->base, X->sample_stride, X->col_stride, X->batch_stride};
  if (cdiv(n_samples, 4) > 65535 || cdiv(n_targets, 4) > 65535) return fail(MVB_E_UNSUPPORTED, "dense_predict_batched: too many samples / targets");
  mvb::db::predict_kernel<<<dim3(cdiv(batch, 32), cdiv(n_samples, 4), cdiv(n_targets, 4)), dim3(32, 4), 0, st>>>(
      xv, n_samples, n_cols, batch, (const float*)coef, (const float*)intercept, n_targets, (float*)yhat, out_sample_stride,
      out_target_stride, out_batch_stride);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_rank(const void* x, int64_t R, int64_t K, int dtype, void* out, void* stream) {
  if (!x || !out || R < 1 || K < 1) return fail(MVB_E_ARG, "rank: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(K, 128);
  if (dtype == MVB_F32) {
    if (R <= 64) mvb::rank_kernel<float, 64><<<grid, 128, 0, st>>>((const float*)x, R, K, (float*)out);
    else mvb::rank_kernel<float, 0><<<grid, 128, 0, st>>>((const float*)x, R, K, (float*)out);
  } else if (dtype == MVB_F64) {
    if (R <= 64) mvb::rank_kernel<double, 64><<<grid, 128, 0, st>>>((const double*)x, R, K, (double*)out);
    else mvb::rank_kernel<double, 0><<<grid, 128, 0, st>>>((const double*)x, R, K, (double*)out);
  } else return fail(MVB_E_UNSUPPORTED, "rank: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_score(const void* y, const void* yhat, const void* y_b, int64_t R, int64_t K, int dtype, void* r2_out,
              void* pearson_out, void* stream) {
  if (!y || !yhat || R < 1 || K < 1 || (!r2_out && !pearson_out)) return fail(MVB_E_ARG, "score: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = cdiv(K, 128);
  // vectorised warp-shuffle kernel when rows are 16-byte tileable; the scalar kernel keeps ragged shapes
  const bool vec32 = K % 4 == 0 && (((uintptr_t)y | (uintptr_t)yhat | (uintptr_t)y_b) & 15) == 0;
  const bool vec64 = K % 2 == 0 && (((uintptr_t)y | (uintptr_t)yhat | (uintptr_t)y_b) & 15) == 0;
  if (dtype == MVB_F32) {
    // small problems: four interleaved row slices per warp so that there are enough warps to fill the SMs
    if (vec32 && K / 4 < (int64_t)148 * 2048) mvb::score_vec_kernel<float, 4, 4><<<cdiv(K, 256), 256, 0, st>>>((const float*)y, (const float*)yhat, (const float*)y_b, R, K, (float*)r2_out, (float*)pearson_out);
    else if (vec32) mvb::score_vec_kernel<float, 1, 4><<<cdiv(K, 1024), 256, 0, st>>>((const float*)y, (const float*)yhat, (const float*)y_b, R, K, (float*)r2_out, (float*)pearson_out);
    else if (R <= 32) mvb::score_kernel<float, 32><<<grid, 128, 0, st>>>((const float*)y, (const float*)yhat, (const float*)y_b, R, K, (float*)r2_out, (float*)pearson_out);
    else mvb::score_kernel<float, 0><<<grid, 128, 0, st>>>((const float*)y, (const float*)yhat, (const float*)y_b, R, K, (float*)r2_out, (float*)pearson_out);
  } else if (dtype == MVB_F64) {
    if (vec64 && K / 2 < (int64_t)148 * 2048) mvb::score_vec_kernel<double, 4, 4><<<cdiv(K, 128), 256, 0, st>>>((const double*)y, (const double*)yhat, (const double*)y_b, R, K, (double*)r2_out, (double*)pearson_out);
    else if (vec64) mvb::score_vec_kernel<double, 1, 4><<<cdiv(K, 512), 256, 0, st>>>((const double*)y, (const double*)yhat, (const double*)y_b, R, K, (double*)r2_out, (double*)pearson_out);
    else if (R <= 32) mvb::score_kernel<double, 32><<<grid, 128, 0, st>>>((const double*)y, (const double*)yhat, (const double*)y_b, R, K, (double*)r2_out, (double*)pearson_out);
    else mvb::score_kernel<double, 0><<<grid, 128, 0, st>>>((const double*)y, (const double*)yhat, (const double*)y_b, R, K, (double*)r2_out, (double*)pearson_out);
  }
  else return fail(MVB_E_UNSUPPORTED, "score: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

// ---- ReceptiveField pieces ------------------------------------------------------------------------
int mvb_rf_assemble(const void* G, const void* Xc, int n_trials, int n_feat, int n_lag, int n_time, int s_min, int s_max,
                    int edge_correction, int dtype, void* cov, void* stream) {
  if (!G || !Xc || !cov || n_trials < 1 || n_feat < 1 || n_lag < 1 || n_time < 1 || s_max - s_min != n_lag)
    return fail(MVB_E_ARG, "rf_assemble: bad argument");
  const int L = (edge_correction && s_min < 0) ? -s_min : 0;
  if (L > n_time) return fail(MVB_E_ARG, "rf_assemble: %d negative lags need at least as many samples (%d)", L, n_time);
  const int Lh = L < n_lag ? L : n_lag;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)n_trials * n_feat * n_lag * n_feat * n_lag;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  if (dtype == MVB_F32) mvb::rf::assemble_kernel<float><<<grid, 256, 0, st>>>((const float*)G, (const float*)Xc, n_trials, n_feat, n_lag, n_time, L, Lh, (float*)cov);
  else if (dtype == MVB_F64) mvb::rf::assemble_kernel<double><<<grid, 256, 0, st>>>((const double*)G, (const double*)Xc, n_trials, n_feat, n_lag, n_time, L, Lh, (double*)cov);
  else return fail(MVB_E_UNSUPPORTED, "rf_assemble: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_sum_slabs(const void* A, int64_t slab_elems, const int32_t* idx, int n_idx, int dtype, void* out, void* stream) {
  if (!A || !idx || !out || slab_elems < 1 || n_idx < 0) return fail(MVB_E_ARG, "sum_slabs: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>((slab_elems + 255) / 256, 148 * 16);
  if (dtype == MVB_F32) mvb::rf::sum_slabs_kernel<float><<<grid, 256, 0, st>>>((const float*)A, slab_elems, idx, n_idx, (float*)out);
  else if (dtype == MVB_F64) mvb::rf::sum_slabs_kernel<double><<<grid, 256, 0, st>>>((const double*)A, slab_elems, idx, n_idx, (double*)out);
  else return fail(MVB_E_UNSUPPORTED, "sum_slabs: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_penalised_solve(const void* G, const void* reg, const void* alphas, int n_alphas, const void* rhs, int p, int n_rhs,
                        int dtype, void* mats, void* sol, int* info, void* stream) {
  if (!G || !reg || !alphas || !rhs || !mats || !sol || !info || n_alphas < 1 || p < 1 || n_rhs < 1)
    return fail(MVB_E_ARG, "penalised_solve: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t pp = (int64_t)p * p;
  const size_t es = dtype == MVB_F32 ? 4 : 8;
  const int grid = (int)std::min<int64_t>((pp * n_alphas + 255) / 256, 148 * 16);
  for (int s = 0; s < n_alphas; ++s) {      // every system starts from the same right-hand sides
    cudaError_t e = cudaMemcpyAsync((char*)sol + (size_t)s * p * n_rhs * es, rhs, (size_t)p * n_rhs * es, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return fail(MVB_E_CUDA, "penalised_solve: %s", cudaGetErrorString(e));
  }
  if (dtype == MVB_F32) {
    mvb::rf::penalised_kernel<float><<<grid, 256, 0, st>>>((const float*)G, (const float*)reg, (const float*)alphas, n_alphas, pp, (float*)mats);
    mvb::rf::lu_solve_kernel<float><<<n_alphas, mvb::rf::LU_THREADS, 0, st>>>((float*)mats, (float*)sol, p, n_rhs, info);
  } else if (dtype == MVB_F64) {
    mvb::rf::penalised_kernel<double><<<grid, 256, 0, st>>>((const double*)G, (const double*)reg, (const double*)alphas, n_alphas, pp, (double*)mats);
    mvb::rf::lu_solve_kernel<double><<<n_alphas, mvb::rf::LU_THREADS, 0, st>>>((double*)mats, (double*)sol, p, n_rhs, info);
  } else return fail(MVB_E_UNSUPPORTED, "penalised_solve: dtype %d", dtype);
  MVB_LAUNCHED();
  MVB_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MVB_E_CUDA, "penalised_solve: %s", cudaGetErrorString(e));
  return MVB_OK;
}

// ---- classification metrics ------------------------------------------------------------------------
int mvb_accuracy(const void* y, const void* yhat, int64_t R, int64_t K, int dtype, void* out, void* stream) {
  if (!y || !yhat || !out || R < 1 || K < 1) return fail(MVB_E_ARG, "accuracy: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MVB_F32) mvb::accuracy_kernel<float><<<cdiv(K, 128), 128, 0, st>>>((const float*)y, (const float*)yhat, R, K, (float*)out);
  else if (dtype == MVB_F64) mvb::accuracy_kernel<double><<<cdiv(K, 128), 128, 0, st>>>((const double*)y, (const double*)yhat, R, K, (double*)out);
  else return fail(MVB_E_UNSUPPORTED, "accuracy: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

size_t mvb_pairwise_coupling_workspace_bytes(int64_t n_samples, int n_classes, int max_iter) {
  return (size_t)(2 * n_samples * n_classes + max_iter) * sizeof(double) + 256;
}

int mvb_pairwise_coupling(const void* pairwise, int64_t n_samples, int n_classes, int max_iter, double eps, double tol, int dtype,
                          void* out, void* ws, size_t ws_bytes, void* stream) {
  if (!pairwise || !out || !ws || n_samples < 1 || n_classes < 2 || max_iter < 1) return fail(MVB_E_ARG, "pairwise_coupling: bad argument");
  if (n_classes > mvb::COUPLING_MAX_K) return fail(MVB_E_UNSUPPORTED, "pairwise_coupling: %d classes (at most %d)", n_classes, mvb::COUPLING_MAX_K);
  if (ws_bytes < mvb_pairwise_coupling_workspace_bytes(n_samples, n_classes, max_iter))
    return fail(MVB_E_WORKSPACE, "pairwise_coupling: workspace %zu too small", ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = n_samples * n_classes;
  double* buf[2] = {(double*)ws, (double*)ws + total};
  double* maxdiff = (double*)ws + 2 * total;
  MVB_CUDA(cudaMemsetAsync(maxdiff, 0, (size_t)max_iter * sizeof(double), st));
  const int grid = cdiv(n_samples, 64);
  for (int it = 0; it < max_iter; ++it) {
    if (dtype == MVB_F32) mvb::coupling_step_kernel<float><<<grid, 64, 0, st>>>((const float*)pairwise, (int)n_samples, n_classes, eps, tol, it, buf[it & 1], buf[(it + 1) & 1], maxdiff);
    else if (dtype == MVB_F64) mvb::coupling_step_kernel<double><<<grid, 64, 0, st>>>((const double*)pairwise, (int)n_samples, n_classes, eps, tol, it, buf[it & 1], buf[(it + 1) & 1], maxdiff);
    else return fail(MVB_E_UNSUPPORTED, "pairwise_coupling: dtype %d", dtype);
    MVB_LAUNCHED();
  }
  if (dtype == MVB_F32) mvb::coupling_finish_kernel<float><<<grid_for(total), 256, 0, st>>>(buf[0], buf[1], maxdiff, max_iter, tol, total, (float*)out);
  else mvb::coupling_finish_kernel<double><<<grid_for(total), 256, 0, st>>>(buf[0], buf[1], maxdiff, max_iter, tol, total, (double*)out);
  MVB_LAUNCHED();
  return MVB_OK;
}

int mvb_roc_auc(const void* positive, const void* score, int64_t R, int64_t K, int dtype, void* out, void* stream) {
  if (!positive || !score || !out || R < 1 || K < 1) return fail(MVB_E_ARG, "roc_auc: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MVB_F32) mvb::auc_kernel<float><<<cdiv(K, 128), 128, 0, st>>>((const float*)positive, (const float*)score, R, K, (float*)out);
  else if (dtype == MVB_F64) mvb::auc_kernel<double><<<cdiv(K, 128), 128, 0, st>>>((const double*)positive, (const double*)score, R, K, (double*)out);
  else return fail(MVB_E_UNSUPPORTED, "roc_auc: dtype %d", dtype);
  MVB_LAUNCHED();
  return MVB_OK;
}

}  // extern "C"
