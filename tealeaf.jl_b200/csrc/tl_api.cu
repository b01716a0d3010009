// tl_api.cu -- context, C-ABI entry points (include/tealeaf_b200.h) and the whole-solve
// drivers of libtealeaf_b200.so.  Built for sm_100a only; there is no CPU fallback: every
// entry point fails with TL_ERR_NO_DEVICE / TL_ERR_CUDA when no usable GPU is present.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <climits>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <unistd.h>
#ifdef TL_WITH_NCCL
#include <nccl.h>   // types only: the library is resolved at run time (see NcclApi)
#include <dlfcn.h>
#endif

#include "../../include/tealeaf_b200.h"
#include "tl_device.cuh"
#include "tl_kernels_basic.cuh"
#include "tl_kernels_fused.cuh"
#include "tl_kernels_ring.cuh"
#include "tl_kernels_persist.cuh"
#include "tl_kernels_tma.cuh"
#include "tl_eigen.h"

#define TL_MAX_GRID 4096
#define TL_ERROR_START 1e10       // src/kernels.jl:7
#define TL_ERROR_SWITCH_MAX 1.0   // src/kernels.jl:8
#define TL_CGEIGENITERS 20        // src/solvers/Cheby.jl:7

// ping-pong partners of p, u, sd; the working buffers of the matrix-powers PPCG groups
// (PpcgDkParams: sd wa/wb, the second group-input copy of r, r's working buffer)
enum { B_P1 = TL_NUM_FIELDS, B_U1, B_SD1, B_SD2, B_SD3, B_R1, B_R2, B_COUNT };

#ifdef TL_WITH_NCCL
// NCCL is bound lazily with dlopen instead of at link time: a process that also imports torch
// must end up with ONE libnccl.so.2 (torch bundles a newer one than the system's), so an
// already-loaded copy is preferred (RTLD_NOLOAD) and the system copy is the fallback.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi &nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString;
  return api;
}
#endif

struct CommBlob {  // what tl_comm_export hands to the other ranks
  cudaIpcMemHandle_t handle;
  int nx, ny, hd, pitch;
  long long buf_offset[B_COUNT];  // byte offset of interior cell (0,0) of each buffer in the slab
  long long mail_offset;          // byte offset of the tile's mailbox (tl_tile_exchange)
  int rank, device;
  long long pid;                  // exporting process: tiles of the SAME process are wired by plain pointer
  unsigned long long slab_ptr;    // (CUDA-IPC handles cannot be opened by the process that made them)
};

struct Multi;   // tl_multi.inl: a context that spans several GPUs of this process

struct tl_ctx {
  Multi *multi = nullptr;   // non-null: this context is the single-Chunk face of px x py tile contexts (tl_create_multi)
  Geo g{};
  int max_iters = 0, device = 0;
  int rank = 0, px = 1, py = 1, cx = 0, cy = 0, nranks = 1;
  size_t rows = 0, buf_doubles = 0, hist_len = 0;   // hist_len: doubles in each of hist_rr / hist_pw / ch_alphas / ch_betas
  char *slab = nullptr;
  size_t slab_bytes = 0;
  double *buf[B_COUNT]{};   // interior-origin pointers
  int p_cur = 0, u_cur = 0, sd_cur = 0;
  SolveState *st = nullptr;
  double *hist_rr = nullptr, *hist_pw = nullptr, *ch_alphas = nullptr, *ch_betas = nullptr, *partials = nullptr;
  SolveState *h_st = nullptr;   // pinned, 2 polling slots
  double *h_scal = nullptr;     // pinned scratch
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[2]{}, ev_start = nullptr, ev_stop = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_phase = nullptr;
  bool phase_marked = false;
  double last_cg_ms = 0.0;      // device time of the CG phase (preamble + CG presteps + flush) of the last solve: tl_get_option("last_cg_phase_ms")
  int num_sms = 148;
  // tuning
  int blocks_per_sm = 2, pw_blocks_per_sm = 4, chunk_rows = -1, pw_chunk_rows = -1, graph_iters = 8, use_graph = 1;
  int ring_stages = -1;  // -1: auto; 3/4/6: cp.async shared-memory ring depth
  int ring_eff = 0;      // the flavour in use
  int hint_keep = TL_HINT_NORMAL, hint_stream = TL_HINT_NORMAL, b_reverse = 1;   // L2 reuse between CG kernels A and B
  double l2_persist_mb = 0.0, l2_hit_scale = 1.0;
  int l2_persist_field = TL_R;
  Tiling tiling{}, pw_tiling{};   // stencil kernels / pointwise kernels
  int fused_grid = 0, pw_grid = 0, basic_grid = 0;
  DkExt dk_ext{};                 // k_ppcg_inner_dk: the extension warps appended to `tiling`
  int dk_grid = 0, dk_k = 1, dk_rows_per_chunk = 0;
  int ppcg_depth_k = 0;           // option "ppcg_halo_depth": 0 = auto (halo_depth)
  cudaGraphExec_t g_cg = nullptr, g_cheby = nullptr, g_ppcg = nullptr, g_jacobi = nullptr;
  int g_cg_iters = 0, g_cheby_iters = 0, g_ppcg_iters = 0, g_ppcg_inner = 0, g_ppcg_k = 0, g_ppcg_pairs = 0, g_jacobi_iters = 0;
  long long launches = 0;
  // peers (tile-internal sides): 0 left, 1 right, 2 bottom, 3 top
  int nbr_rank[4] = {-1, -1, -1, -1};
  void *peer_slab[4]{};
  CommBlob peer_blob[4]{};
  CommBlob rank_blob[TL_MAX_RANKS]{};
  void *rank_slab[TL_MAX_RANKS]{};   // every other tile's slab, CUDA-IPC mapped (mailboxes; the neighbours' fields)
  bool rank_ipc[TL_MAX_RANKS]{};     // true: mapped with cudaIpcOpenMemHandle (to be closed); false: same-process pointer
  bool comm_ready = false;
  int prof = 0;             // kernel-boundary micro-profile (option prof, tl_get_option prof_*)
  int use_pdl = 0;          // programmatic dependent launch between the kernels of the iteration loops: measured
                            // SLOWER (profiles/r01d_pdl_sweep.log: 1024^2 CG 24.7 -> 27.6 us/iteration), kept as an option
  int pair_tiled = 1;       // 1: the pair kernels also run on tiles (depth-2 halos, one rendezvous per two iterations / inner steps)
  int ppcg_pair = 1;        // 1: PPCG inner steps run two per pass (k_ppcg_pair_ring; an odd count ends with one single step)
  int cheby_pair = 1;       // 1: reduction-free Chebyshev iterations run two per pass (k_cheby_pair_ring)
  int pair_rows = 32;       // rows per warp task of the pair kernel (two redundant rows per task)
  int pair_stages = 4;      // cp.async ring depth of the pair kernels: 4 (default) or 5, two CTAs per SM either way
  Tiling pair_tiling{};
  int pair_grid = 0;
  cudaGraphExec_t g_cheby2 = nullptr;
  int g_cheby2_iters = 0;
  int balanced_tiling = 1;  // mid-size tiles: chunk length chosen so that every SM holds exactly two CTAs (compute_tiling)
  int a_tma = 0;            // 1: kernel A's ring is filled by TMA (cp.async.bulk.tensor boxes + mbarriers, tl_kernels_tma.cuh); measured, off
  CUtensorMap tma_maps[TMA_NMAPS];
  bool tma_ready = false;
  int b_ring = 0;           // kernel B flavour: 0 = register batches of 4 rows; 6 / 8 = cp.async ring of that depth (4 / 3 CTAs per SM)
  int lazy_heavy_ctas = 2;  // CTAs per SM of the u-updating launch of the lazy loop at ring depth 3: 2 (116 registers, no spill, 164 KB
                            // carve-out; measured 0.5-1 % faster per iteration at 4096^2) or 3 (80 registers, all-shared carve-out)
  int cg_seq = 0;           // position of the next CG iteration in its chunk (lazy-u loop: which of the two kernels)
  int cg_lazy_u = 1;        // 1: the CG loop advances u every second iteration with both pending updates (TL_U_LAZY, tl_kernels_ring.cuh)
  int cg_persist = 0;       // 1: the CG loop of a single tile runs as ONE persistent cooperative kernel (tl_kernels_persist.cuh)
  PersistSync *psync = nullptr;
  int comm_fused = 1;       // 1: halo pushes + mailbox allreduce inside the kernels; 0: halo-pull kernels + NCCL
  int xchg_deferred = 0;    // 1: split exchange -- kernels post their packets in the tail, the NEXT kernel collects them at its entry
  CommDev *d_comm = nullptr;   // device copy of the mailbox table (in the slab)
  MailSlot *mail = nullptr;
#ifdef TL_WITH_NCCL
  ncclComm_t nccl = nullptr;
#endif
  std::string err;
};

static int tl_fail(tl_ctx *c, int code, const char *fmt, ...) {
  char b[512];
  va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
  if (c) c->err = b;
  return code;
}
#define CU(c, call)                                                                               \
  do { cudaError_t _e = (call);                                                                   \
       if (_e != cudaSuccess) return tl_fail((c), TL_ERR_CUDA, "%s failed: %s (%s:%d)", #call,    \
                                             cudaGetErrorString(_e), __FILE__, __LINE__); } while (0)
#define CHECK_LAUNCH(c) CU(c, cudaGetLastError())
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

// Launch of a hot-loop kernel.  With use_pdl the launch carries the programmatic-stream-
// serialization attribute (a programmatic edge when captured into a graph): the kernel may become
// resident while its predecessor drains, and waits in tl_pdl_entry() (griddepcontrol.wait) until
// the predecessor has completed and flushed -- the launch latency between the dependent kernels of
// an iteration disappears from the critical path.
template <typename P>
static cudaError_t tl_launch(tl_ctx *c, void (*kern)(const P), int grid, int block, size_t smem, const P &params);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember which devices a
// kernel instantiation has been prepared on (one static mask per instantiation).
// Experiment, off (process-wide, as kernel attributes are): TEALEAF_B200_CARVEOUT=1 makes every loop kernel ask for the
// all-shared L1 / shared-memory split, so that consecutive kernels never make the SMs reconfigure.  Measured on 2 GPUs
// (profiles/r02i_carveout_acquire_ab_n2.jsonl): the launch gap does not shrink (3.2 vs 3.0 us) and the kernel bodies slow down
// by 10-16 % (28 KB of L1 instead of 60 KB for the edge and prologue loads), so each kernel keeps its own split.
static const bool g_same_carveout = getenv("TEALEAF_B200_CARVEOUT") && atoi(getenv("TEALEAF_B200_CARVEOUT")) == 1;
template <typename P>
static int tl_prepare_smem(tl_ctx *c, void (*kern)(const P), int smem, std::atomic<unsigned long long> *device_mask) {
  const unsigned long long bit = 1ull << (c->device & 63);
  if (!(device_mask->load() & bit)) {   // tiles of one process prepare from several service threads: idempotent, the mask is atomic
    if (smem > 0) CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // every kernel of the iteration loops asks for the SAME L1 / shared-memory split (all shared): two consecutive
    // kernels with different carve-outs make the SMs reconfigure between them, which lengthens the launch gap
    if (g_same_carveout) CU(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    device_mask->fetch_or(bit);
  }
  return TL_OK;
}

template <typename P>
static cudaError_t tl_launch(tl_ctx *c, void (*kern)(const P), int grid, int block, size_t smem, const P &params) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = c->use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, params);
}

static double *field_ptr(tl_ctx *c, int f) {
  if (f == TL_P) return c->buf[c->p_cur ? B_P1 : TL_P];
  if (f == TL_U) return c->buf[c->u_cur ? B_U1 : TL_U];
  if (f == TL_SD) return c->buf[c->sd_cur ? B_SD1 : TL_SD];
  return c->buf[f];
}

#define TL_STENCIL_CHUNK_ROWS 12

static void make_tiling(const tl_ctx *c, int blocks_per_sm, int chunk_rows, Tiling *tout, int *grid) {
  const Geo &g = c->g;
  Tiling t;
  t.nstrips = (g.nx + TL_STRIP - 1) / TL_STRIP;
  const int wpb = TL_FUSED_THREADS / 32;
  const int target_warps = c->num_sms * blocks_per_sm * wpb;   // one co-resident wave
  int nchunks = std::max(1, target_warps / t.nstrips);
  if (chunk_rows > 0) nchunks = (g.ny + chunk_rows - 1) / chunk_rows;
  nchunks = std::min(nchunks, g.ny);
  nchunks = std::min(nchunks, std::max(1, TL_MAX_GRID * wpb / t.nstrips));  // bounded partials array
  t.rows_per_chunk = (g.ny + nchunks - 1) / nchunks;
  t.nchunks = (g.ny + t.rows_per_chunk - 1) / t.rows_per_chunk;
  *tout = t;
  *grid = (t.nstrips * t.nchunks + wpb - 1) / wpb;
}

static void compute_tiling(tl_ctx *c) {
  const Geo &g = c->g;
  // Measured on B200 (profiles/r01_tuning.md): short row chunks issued as several waves of
  // memory-contiguous CTAs beat one co-resident wave; the cp.async ring wins once chunks are short.
  // Depth 3 at 3 CTAs/SM is best up to ~4096^2 tiles, depth 4 at 2 CTAs/SM beyond.
  const long cells_tile = (long)g.nx * g.ny;
  c->ring_eff = c->ring_stages >= 0 ? c->ring_stages : (cells_tile >= (long)8192 * 8192 ? 4 : 3);
  const int bps = (c->ring_eff == 3) ? 3 : (c->ring_eff == 4) ? 2 : 1;   // the kernels' launch bounds
  // rows per warp: 12 (stencil) / 16 (pointwise) once the mesh fills the machine (8 until the lazy-u CG loop: with it
  // whole solves run 2-3 % faster at 11-13 rows at 4096^2 and 2048^2, profiles/r02q_chunk_rows_sweep.log); small meshes
  // get shorter chunks so that at least half a wave of warps exists (profiles/r01c_small_mesh_sweep.log)
  const long warp_rows = (long)((g.nx + TL_STRIP - 1) / TL_STRIP) * g.ny;
  int cr = (int)std::min<long>(TL_STENCIL_CHUNK_ROWS, std::max<long>(1, warp_rows / ((long)c->num_sms * bps * 4)));
  int pcr = (int)std::min<long>(16, std::max<long>(1, warp_rows / ((long)c->num_sms * c->pw_blocks_per_sm * 4)));
  if (pcr >= 4) pcr &= ~3;   // the pointwise kernels are unrolled by 4 rows
  // Mid-size tiles (what a 4096^2 mesh becomes on 4-8 GPUs: 1024^2 ... 2048 x 1024 cells) are only one
  // or two waves of CTAs, so the wave quantisation decides: 512 CTAs on 444 slots run as two waves, and
  // 412 CTAs leave some SMs with 3 and others with 2.  Measured (profiles/r01e_tile_sweep.log): exactly
  // TWO CTAs on every SM, with chunks as long as that takes (up to 16 rows), is fastest there
  // (2048 x 1024: kernel A 32.8 -> 28.3 us, PPCG inner 26.6 -> 22.6 us, B 11.3 -> 10.3 us).
  if (c->balanced_tiling) {
    const int wpb = TL_FUSED_THREADS / 32, nstrips = (g.nx + TL_STRIP - 1) / TL_STRIP;
    const long slots = 2L * c->num_sms;
    auto balanced = [&](int quantum) -> int {
      long r = std::max<long>(1, (warp_rows + slots * wpb - 1) / (slots * wpb));
      if (quantum > 1 && r >= quantum - 1) r = (r + quantum - 1) / quantum * quantum;
      for (; r <= 16; r += (r >= quantum ? quantum : 1)) {
        const long nchunks = (g.ny + r - 1) / r;
        if ((nstrips * nchunks + wpb - 1) / wpb <= slots) return (int)r;
      }
      return 0;   // the tile is several waves whatever the chunk length: keep the streaming defaults
    };
    if (const int r = balanced(1)) cr = r;
    if (const int r = balanced(4)) pcr = r;
  }
  if (c->chunk_rows >= 0) cr = c->chunk_rows;
  if (c->pw_chunk_rows >= 0) pcr = c->pw_chunk_rows;
  make_tiling(c, bps, cr, &c->tiling, &c->fused_grid);
  make_tiling(c, c->pw_blocks_per_sm, pcr, &c->pw_tiling, &c->pw_grid);
  const long cells = (long)(g.nx + 2 * g.hd) * (g.ny + 2 * g.hd);
  long nb = (cells + TL_BASIC_THREADS - 1) / TL_BASIC_THREADS;
  c->basic_grid = (int)std::max(1L, std::min<long>(nb, (long)c->num_sms * 8));
  {  // the two-iteration Chebyshev kernel: 60 owned columns per warp, long chunks (2 redundant rows each)
    const int wpb = TL_FUSED_THREADS / 32;
    Tiling t;
    t.nstrips = (g.nx + TL_PAIR_OWN - 1) / TL_PAIR_OWN;
    const long wr = (long)t.nstrips * g.ny;
    long r = std::min<long>(c->pair_rows, std::max<long>(8, wr / (2L * c->num_sms * wpb)));   // small tiles: >= 2 CTAs per SM first
    r = std::max<long>(r, ((long)t.nstrips * g.ny + (long)TL_MAX_GRID * wpb - 1) / ((long)TL_MAX_GRID * wpb));   // bounded partials array
    r = std::min<long>(std::max<long>(r, 1), g.ny);
    for (;; r++) {   // the rounding of the bound above can leave a few CTAs too many
      t.rows_per_chunk = (int)r;
      t.nchunks = (g.ny + t.rows_per_chunk - 1) / t.rows_per_chunk;
      c->pair_grid = (t.nstrips * t.nchunks + wpb - 1) / wpb;
      if (c->pair_grid <= TL_MAX_GRID || r >= g.ny) break;
    }
    c->pair_tiling = t;
  }
}

// Decomposition of the tile extended by k-1 cells towards its neighbour tiles (k_ppcg_inner_dk):
// the plain kernels' warps for the interior plus appended warps for the extension
// (tl_march_setup_ext).
static void compute_dk_tiling(tl_ctx *c, int k) {
  const Geo &g = c->g;
  const int wpb = TL_FUSED_THREADS / 32;
  const Tiling &t = c->tiling;
  DkExt x;
  x.emax = k - 1;
  x.left = !(g.phys & TL_PHYS_LEFT);
  x.right = !(g.phys & TL_PHYS_RIGHT) && (t.nstrips * TL_STRIP < g.nx + x.emax);
  x.bottom = !(g.phys & TL_PHYS_BOTTOM);
  x.top = !(g.phys & TL_PHYS_TOP);
  const int nall = t.nstrips + x.left + x.right;
  const int warps = t.nstrips * t.nchunks + nall * (x.bottom + x.top) + t.nchunks * (x.left + x.right);
  c->dk_k = k;
  c->dk_ext = x;
  c->dk_rows_per_chunk = t.rows_per_chunk;
  c->dk_grid = (warps + wpb - 1) / wpb;   // <= 4 * TL_MAX_GRID partial slots exist and one value is summed
}

// Optional L2 residency: marks (part of) one field as persisting in the 126 MB L2 through the
// stream's access-policy window, so that its HBM traffic disappears when the mesh is small
// enough (4096^2: one field is 134.5 MB).  Captured graphs inherit the window.
static int apply_l2_policy(tl_ctx *c) {
  cudaStreamAttrValue attr;
  memset(&attr, 0, sizeof attr);
  if (c->l2_persist_mb <= 0.0) {
    attr.accessPolicyWindow.num_bytes = 0;
    CU(c, cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    cudaCtxResetPersistingL2Cache();
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    return TL_OK;
  }
  cudaDeviceProp prop;
  CU(c, cudaGetDeviceProperties(&prop, c->device));
  size_t want = (size_t)(c->l2_persist_mb * 1e6);
  const size_t carve = std::min(want, (size_t)prop.persistingL2CacheMaxSize);
  CU(c, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve));
  const Geo &g = c->g;
  char *base = (char *)(c->buf[c->l2_persist_field] - (size_t)g.hd * g.pitch - TL_XPAD);
  size_t bytes = c->rows * (size_t)g.pitch * sizeof(double);
  bytes = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
  attr.accessPolicyWindow.base_ptr = base;
  attr.accessPolicyWindow.num_bytes = bytes;
  attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve * c->l2_hit_scale / (double)bytes);
  attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  CU(c, cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return TL_OK;
}

static void destroy_graphs(tl_ctx *c) {
  if (c->g_cg) { cudaGraphExecDestroy(c->g_cg); c->g_cg = nullptr; }
  if (c->g_cheby) { cudaGraphExecDestroy(c->g_cheby); c->g_cheby = nullptr; }
  if (c->g_cheby2) { cudaGraphExecDestroy(c->g_cheby2); c->g_cheby2 = nullptr; }
  if (c->g_ppcg) { cudaGraphExecDestroy(c->g_ppcg); c->g_ppcg = nullptr; }
  if (c->g_jacobi) { cudaGraphExecDestroy(c->g_jacobi); c->g_jacobi = nullptr; }
}


// ---- multi-GPU-in-one-process contexts (tl_multi.inl): every entry point forwards to the tiles ----
template <typename F> static int multi_all(tl_ctx *c, F fn);                         // fn(tile, idx) on every tile, concurrently
template <typename F> static int multi_scalar(tl_ctx *c, double *out, F fn);         // fn(tile, &v): the all-tiles value
template <typename F> static int multi_max(tl_ctx *c, double *out, F fn);            // fn(tile, &v): max over tiles
template <typename F> static int multi_min(tl_ctx *c, double *out, F fn);            // fn(tile, &v): min over tiles
template <typename F> static int multi_solve(tl_ctx *c, tl_solve_info *info, F fn);  // fn(tile, &info, idx)
static void multi_destroy(tl_ctx *c);
static int multi_set_field(tl_ctx *c, int field, const double *host, long ld);
static int multi_get_field(tl_ctx *c, int field, double *host, long ld);
static int multi_field_summary(tl_ctx *c, double cell_volume, double *vol, double *mass, double *ie, double *temp);
static int multi_launch_count(tl_ctx *c, long long *count);
static int multi_tile_offset(const tl_ctx *c, int idx, int *x0, int *y0);
#define TL_IS_MULTI(c) ((c) && (c)->multi)

extern "C" int tl_abi_version(void) { return TL_ABI_VERSION; }

// why the last tl_create* call of the process failed (there is no context to ask then); tiles are created from several threads
static std::string g_create_error;
static std::mutex g_create_error_mu;
static void set_create_error(const std::string &msg) {
  std::lock_guard<std::mutex> lk(g_create_error_mu);
  g_create_error = msg;
}
extern "C" const char *tl_last_error(const tl_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

// How the CG loop's kernel A advances u: every second launch (TL_U_LAZY, the default) unless a flavour without that
// mode is selected (TMA ring, persistent kernel)
static int cg_u_mode(const tl_ctx *c) {
  // chunks of an even number of iterations: the two kernels of the lazy loop alternate by position in the chunk
  return (c->cg_lazy_u && !c->a_tma && !(c->cg_persist && c->nranks == 1) && c->graph_iters % 2 == 0) ? TL_U_LAZY : TL_U_EVERY;
}

extern "C" int tl_set_option(tl_ctx *c, const char *name, double value);
extern "C" void tl_destroy(tl_ctx *c);

__global__ void k_zero_words(unsigned long long *p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0ull;
}

extern "C" int tl_create_tile(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int device,
                              int rank, int px, int py) {
  if (!out) return TL_ERR_ARG;
  *out = nullptr;
  if (xcells < 1 || ycells < 1 || halo_depth < 1 || halo_depth > TL_XPAD || max_iters < 1 || px < 1 || py < 1 ||
      rank < 0 || rank >= px * py || px * py > TL_MAX_RANKS)
    return TL_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device >= ndev) {
    cudaGetLastError();
    return TL_ERR_NO_DEVICE;
  }
  tl_ctx *c = new tl_ctx();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return TL_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    delete c;
    return TL_ERR_NO_DEVICE;  // sm_100a-only binary
  }
  c->num_sms = prop.multiProcessorCount;
  c->max_iters = max_iters;
  c->rank = rank; c->px = px; c->py = py; c->cx = rank % px; c->cy = rank / px; c->nranks = px * py;
  Geo &g = c->g;
  g.nx = xcells; g.ny = ycells; g.hd = halo_depth;
  g.pitch = TL_XPAD + ((xcells + halo_depth + 2 + 15) / 16) * 16;
  g.phys = 0;
  if (c->cx == 0) g.phys |= TL_PHYS_LEFT; else c->nbr_rank[0] = rank - 1;
  if (c->cx == px - 1) g.phys |= TL_PHYS_RIGHT; else c->nbr_rank[1] = rank + 1;
  if (c->cy == 0) g.phys |= TL_PHYS_BOTTOM; else c->nbr_rank[2] = rank - px;
  if (c->cy == py - 1) g.phys |= TL_PHYS_TOP; else c->nbr_rank[3] = rank + px;
  c->rows = (size_t)ycells + 2 * halo_depth;
  c->buf_doubles = ((c->rows * g.pitch + 31) / 32) * 32;

  const size_t hist = ((size_t)max_iters + 8 + 31) / 32 * 32;
  c->hist_len = hist;
  size_t bytes = (size_t)B_COUNT * c->buf_doubles * sizeof(double);
  const size_t off_state = bytes; bytes += 4096;
  const size_t off_hist = bytes; bytes += 4 * hist * sizeof(double);
  const size_t off_part = bytes; bytes += 4 * TL_MAX_GRID * sizeof(double);
  const size_t off_mail = bytes; bytes += 2 * TL_MAX_RANKS * sizeof(MailSlot);
  const size_t off_comm = bytes; bytes += (sizeof(CommDev) + 255) / 256 * 256;
  const size_t off_psync = bytes; bytes += (sizeof(PersistSync) + 255) / 256 * 256;
  c->slab_bytes = bytes;
  cudaError_t e = cudaMalloc((void **)&c->slab, bytes);
  if (e != cudaSuccess) {
    set_create_error(cudaGetErrorString(e));
    delete c;
    return TL_ERR_CUDA;
  }
  cudaMemset(c->slab, 0, bytes);
  for (int b = 0; b < B_COUNT; b++)
    c->buf[b] = (double *)c->slab + (size_t)b * c->buf_doubles + (size_t)halo_depth * g.pitch + TL_XPAD;
  c->st = (SolveState *)(c->slab + off_state);
  c->hist_rr = (double *)(c->slab + off_hist);
  c->hist_pw = c->hist_rr + hist;
  c->ch_alphas = c->hist_pw + hist;
  c->ch_betas = c->ch_alphas + hist;
  c->partials = (double *)(c->slab + off_part);
  c->mail = (MailSlot *)(c->slab + off_mail);
  c->d_comm = (CommDev *)(c->slab + off_comm);
  c->psync = (PersistSync *)(c->slab + off_psync);
  if (cudaMallocHost((void **)&c->h_st, 2 * sizeof(SolveState)) != cudaSuccess ||
      cudaMallocHost((void **)&c->h_scal, 64 * sizeof(double)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_stop) != cudaSuccess ||
      cudaEventCreate(&c->ev_t0) != cudaSuccess || cudaEventCreate(&c->ev_t1) != cudaSuccess ||
      cudaEventCreate(&c->ev_phase) != cudaSuccess) {
    set_create_error(cudaGetErrorString(cudaGetLastError()));
    tl_destroy(c);
    return TL_ERR_CUDA;
  }
  compute_tiling(c);
  // The control block (solve state, partials, mailboxes, mailbox table) is written once more by ordinary stores of a
  // kernel: cudaMemset may leave its lines in the memory system's cleared-line state, and the first 8-byte packet another
  // tile's kernel stored into such a line was occasionally never seen by the polling tile (tiles sharing one GPU,
  // profiles/r02f_multi_stress_*.log).
  if (!getenv("TL_NO_TOUCH")) {
    const size_t words = (bytes - off_state) / sizeof(unsigned long long);
    k_zero_words<<<64, 256, 0, c->stream>>>((unsigned long long *)(c->slab + off_state), words);
    cudaStreamSynchronize(c->stream);
  }
  cudaDeviceSynchronize();
  *out = c;
  // default options for every context of the process: TEALEAF_B200_OPTS="name=value,name=value"
  // (A/B runs of bench.py and the test-suite without code changes); unknown names are an error
  if (const char *env = getenv("TEALEAF_B200_OPTS")) {
    std::string all(env);
    size_t pos = 0;
    while (pos < all.size()) {
      size_t end = all.find(',', pos);
      if (end == std::string::npos) end = all.size();
      const std::string kv = all.substr(pos, end - pos);
      pos = end + 1;
      const size_t eq = kv.find('=');
      if (kv.empty()) continue;
      if (eq == std::string::npos || tl_set_option(c, kv.substr(0, eq).c_str(), atof(kv.c_str() + eq + 1)) != TL_OK) {
        set_create_error("bad TEALEAF_B200_OPTS entry: " + kv);
        tl_destroy(c);
        *out = nullptr;
        return TL_ERR_ARG;
      }
    }
  }
  return TL_OK;
}

extern "C" int tl_create(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int device) {
  return tl_create_tile(out, xcells, ycells, halo_depth, max_iters, device, 0, 1, 1);
}

extern "C" void tl_destroy(tl_ctx *c) {
  if (TL_IS_MULTI(c)) { multi_destroy(c); return; }
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  destroy_graphs(c);
#ifdef TL_WITH_NCCL
  if (c->nccl && nccl_api().ok) nccl_api().CommDestroy(c->nccl);
#endif
  for (int r = 0; r < TL_MAX_RANKS; r++)
    if (c->rank_slab[r] && c->rank_ipc[r]) cudaIpcCloseMemHandle(c->rank_slab[r]);
  if (c->ev[0]) cudaEventDestroy(c->ev[0]);
  if (c->ev[1]) cudaEventDestroy(c->ev[1]);
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  if (c->ev_stop) cudaEventDestroy(c->ev_stop);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_phase) cudaEventDestroy(c->ev_phase);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->h_st) cudaFreeHost(c->h_st);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->slab) cudaFree(c->slab);
  delete c;
}

__global__ void k_state_prof(SolveState *st, int on);
extern "C" int tl_set_option(tl_ctx *c, const char *name, double value) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_set_option(t, name, value); });
  if (!c || !name) return TL_ERR_ARG;
  const std::string n(name);
  if (n == "blocks_per_sm") c->blocks_per_sm = std::max(1, (int)value);
  else if (n == "pw_blocks_per_sm") c->pw_blocks_per_sm = std::max(1, (int)value);
  else if (n == "ring_stages") {
    const int s = (int)value;
    if (s != -1 && s != 3 && s != 4 && s != 6) return tl_fail(c, TL_ERR_ARG, "ring_stages must be -1, 3, 4 or 6");
    c->ring_stages = s;
  }
  else if (n == "hint_keep") c->hint_keep = std::min(std::max(0, (int)value), 2);
  else if (n == "hint_stream") c->hint_stream = std::min(std::max(0, (int)value), 2);
  else if (n == "b_reverse") c->b_reverse = value != 0.0;
  else if (n == "comm_fused") c->comm_fused = value != 0.0;
  else if (n == "xchg_deferred") {
    c->xchg_deferred = value != 0.0;
    if (c->comm_ready && c->nranks > 1) {   // the kernels read the switch from the device copy of the mailbox table
      cudaSetDevice(c->device);
      cudaStreamSynchronize(c->stream);
      const int flag[2] = {c->xchg_deferred, 0};
      memcpy(c->h_scal, flag, sizeof flag);
      if (cudaMemcpyAsync(&c->d_comm->deferred, c->h_scal, sizeof flag, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess)
        return tl_fail(c, TL_ERR_CUDA, "xchg_deferred: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  else if (n == "use_pdl") c->use_pdl = value != 0.0;   // programmatic dependent launch, released before the kernel tails (tl_pdl_trigger)
  else if (n == "cg_persist") c->cg_persist = value != 0.0;
  else if (n == "cg_lazy_u") c->cg_lazy_u = value != 0.0;
  else if (n == "cg_lazy_heavy_ctas") c->lazy_heavy_ctas = value == 3.0 ? 3 : 2;
  else if (n == "a_tma") {   // 0 off; 1 / 4: TMA ring of 4 row slots; 3: of 3 row slots (two CTAs per SM either way)
    const int d = (int)value;
    if (d != 0 && d != 1 && d != 3 && d != 4) return tl_fail(c, TL_ERR_ARG, "a_tma must be 0, 1, 3 or 4");
    c->a_tma = d == 1 ? 4 : d;
  }
  else if (n == "balanced_tiling") c->balanced_tiling = value != 0.0;
  else if (n == "cheby_pair") c->cheby_pair = value != 0.0;
  else if (n == "ppcg_pair") c->ppcg_pair = value != 0.0;
  else if (n == "pair_tiled") c->pair_tiled = value != 0.0;
  else if (n == "pair_rows") c->pair_rows = std::max(2, (int)value);
  else if (n == "pair_stages") {
    // (3 slots at three CTAs per SM was measured too: 80 registers force spills, 91 -> 144 us per iteration at 4096^2:
    // profiles/r02o_pair_stages_ab.log -- not kept)
    if ((int)value != 4 && (int)value != 5) return tl_fail(c, TL_ERR_ARG, "pair_stages must be 4 or 5");
    c->pair_stages = (int)value;
  }
  else if (n == "b_ring") {
    const int d = (int)value;
    if (d != 0 && d != 6 && d != 8) return tl_fail(c, TL_ERR_ARG, "b_ring must be 0, 6 or 8");
    c->b_ring = d;
  }
  else if (n == "ppcg_halo_depth") {   // exchange depth of the PPCG inner steps on tiles: 0 = auto (halo_depth)
    if (value < 0 || value > c->g.hd) return tl_fail(c, TL_ERR_ARG, "ppcg_halo_depth must be in 0..halo_depth");
    c->ppcg_depth_k = (int)value;
  }
  else if (n == "l2_persist_mb") c->l2_persist_mb = value;
  else if (n == "l2_hit_scale") c->l2_hit_scale = value;
  else if (n == "l2_persist_field") c->l2_persist_field = std::min(std::max(0, (int)value), (int)B_COUNT - 1);
  else if (n == "chunk_rows") c->chunk_rows = std::max(-1, (int)value);       // -1 auto, 0 one wave
  else if (n == "pw_chunk_rows") c->pw_chunk_rows = std::max(-1, (int)value);
  else if (n == "graph_iters") c->graph_iters = std::max(1, (int)value);
  else if (n == "use_graph") c->use_graph = value != 0.0;
  else if (n == "prof") {   // kernel-boundary micro-profile on / off; (re)setting it clears the accumulators
    cudaSetDevice(c->device);
    k_state_prof<<<1, 1, 0, c->stream>>>(c->st, value != 0.0);
    c->prof = value != 0.0;
    return cudaStreamSynchronize(c->stream) == cudaSuccess ? TL_OK : tl_fail(c, TL_ERR_CUDA, "prof: %s", cudaGetErrorString(cudaGetLastError()));
  }
  else return tl_fail(c, TL_ERR_ARG, "unknown option %s", name);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  destroy_graphs(c);
  compute_tiling(c);
  c->dk_grid = 0;   // recomputed by the next depth-k PPCG solve
  // The L2 carve-out calls synchronise the whole DEVICE: only when an L2 option changed (tiles that share a GPU
  // must never device-synchronise while another tile's kernel waits for them in a rendezvous).
  if (n.rfind("l2_", 0) == 0) return apply_l2_policy(c);
  return TL_OK;
}

__global__ void k_state_prof(SolveState *st, int on) {
  st->prof = on;
  st->prof_start = st->prof_prev_end = 0ull;
  for (int q = 0; q < 6; q++) st->prof_acc[q] = 0ull;
}

// Read-back of an option or of a derived quantity (what the tests assert the measured code paths on).
static int cg_a_occupancy(tl_ctx *c, int *out);

extern "C" int tl_get_option(tl_ctx *c, const char *name, double *value) {
  if (TL_IS_MULTI(c)) {
    if (name && strncmp(name, "debug_", 6) == 0) return multi_min(c, value, [&](tl_ctx *t, double *o) { return tl_get_option(t, name, o); });
    return multi_max(c, value, [&](tl_ctx *t, double *o) { return tl_get_option(t, name, o); });
  }
  if (!c || !name || !value) return TL_ERR_ARG;
  const std::string n(name);
  double v;
  if (n == "blocks_per_sm") v = c->blocks_per_sm;
  else if (n == "pw_blocks_per_sm") v = c->pw_blocks_per_sm;
  else if (n == "ring_stages") v = c->ring_stages;
  else if (n == "ring_stages_effective") v = c->ring_eff;
  else if (n == "hint_keep") v = c->hint_keep;
  else if (n == "hint_stream") v = c->hint_stream;
  else if (n == "b_reverse") v = c->b_reverse;
  else if (n == "comm_fused") v = c->comm_fused;
  else if (n == "xchg_deferred") v = c->xchg_deferred;
  else if (n == "use_pdl") v = c->use_pdl;
  else if (n == "cg_persist") v = c->cg_persist;
  else if (n == "cg_lazy_u") v = c->cg_lazy_u;
  else if (n == "cg_lazy_heavy_ctas") v = c->lazy_heavy_ctas;
  else if (n == "cg_u_mode") v = cg_u_mode(c);
  else if (n == "default_chunk_rows") v = TL_STENCIL_CHUNK_ROWS;
  else if (n == "cg_a_blocks_per_sm") {
    int occ = 0;
    CU(c, cudaSetDevice(c->device));
    TRY(cg_a_occupancy(c, &occ));
    v = occ;
  }
  else if (n == "a_tma") v = c->a_tma;
  else if (n == "balanced_tiling") v = c->balanced_tiling;
  else if (n == "cheby_pair") v = c->cheby_pair;
  else if (n == "ppcg_pair") v = c->ppcg_pair;
  else if (n == "pair_tiled") v = c->pair_tiled;
  else if (n == "pair_rows") v = c->pair_rows;
  else if (n == "pair_stages") v = c->pair_stages;
  else if (n == "b_ring") v = c->b_ring;
  else if (n == "ppcg_halo_depth") v = c->ppcg_depth_k;
  else if (n == "l2_persist_mb") v = c->l2_persist_mb;
  else if (n == "l2_hit_scale") v = c->l2_hit_scale;
  else if (n == "l2_persist_field") v = c->l2_persist_field;
  else if (n == "chunk_rows") v = c->chunk_rows;
  else if (n == "pw_chunk_rows") v = c->pw_chunk_rows;
  else if (n == "graph_iters") v = c->graph_iters;
  else if (n == "use_graph") v = c->use_graph;
  // derived (read-only): how the tile is cut
  else if (n == "rows_per_chunk") v = c->tiling.rows_per_chunk;
  else if (n == "pw_rows_per_chunk") v = c->pw_tiling.rows_per_chunk;
  else if (n == "pair_rows_per_chunk") v = c->pair_tiling.rows_per_chunk;
  else if (n == "fused_grid") v = c->fused_grid;
  else if (n == "pw_grid") v = c->pw_grid;
  else if (n == "pair_grid") v = c->pair_grid;
  else if (n == "max_grid") v = TL_MAX_GRID;
  else if (n == "num_sms") v = c->num_sms;
  else if (n == "last_cg_phase_ms") v = c->last_cg_ms;   // device time of the CG phase of the last solve
  else if (n == "debug_comm_table_ok") {   // does the device copy of the mailbox table still hold what tl_comm_connect wrote?
    v = 1.0;
    if (c->nranks > 1 && c->comm_ready) {
      CommDev h;
      CU(c, cudaSetDevice(c->device));
      CU(c, cudaMemcpyAsync(c->h_scal, c->d_comm, sizeof h, cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
      memcpy(&h, c->h_scal, sizeof h);
      if (h.nranks != c->nranks || h.rank != c->rank) v = 0.0;
      for (int r = 0; r < c->nranks; r++) {
        const MailSlot *want = (r == c->rank) ? c->mail : (MailSlot *)((char *)c->rank_slab[r] + c->rank_blob[r].mail_offset);
        if (h.mail[r] != want) v = 0.0;
      }
    }
  }
  else if (n == "prof") v = c->prof;
  else if (n.rfind("prof_", 0) == 0) {   // averages per profiled kernel, microseconds (SolveState::prof_acc)
    SolveState h;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(&h, c->st, sizeof h, cudaMemcpyDeviceToHost));
    const double cnt = (double)h.prof_acc[4], d = cnt > 0 ? cnt * 1e3 : 1.0;
    if (n == "prof_kernels") v = cnt;
    else if (n == "prof_body_us") v = h.prof_acc[0] / d;        // kernel entry -> the last block reaches the tail
    else if (n == "prof_sum_us") v = h.prof_acc[1] / d;         // ticket, partial sums (fence.sys overlapped)
    else if (n == "prof_xchg_us") v = h.prof_acc[2] / d;        // tile exchange: mailbox stores + wait for the slowest tile
    else if (n == "prof_gap_us") v = h.prof_acc[3] / d;         // end of the previous kernel's tail -> this kernel's entry
    else if (n == "prof_fence_us") v = h.prof_acc[5] / d;       // the fence.sys alone
    else return tl_fail(c, TL_ERR_ARG, "unknown option %s", name);
  }
  else return tl_fail(c, TL_ERR_ARG, "unknown option %s", name);
  *value = v;
  return TL_OK;
}

// ---------------------------------------------------------------------------------------
// multi-GPU wiring
// ---------------------------------------------------------------------------------------
extern "C" int tl_comm_blob_size(void) { return (int)sizeof(CommBlob); }

extern "C" int tl_comm_export(tl_ctx *c, void *blob) {
  if (TL_IS_MULTI(c)) return tl_fail(c, TL_ERR_STATE, "tl_comm_export: a tl_create_multi context wires its own tiles");
  if (!c || !blob) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  CommBlob b;
  memset(&b, 0, sizeof b);
  CU(c, cudaIpcGetMemHandle(&b.handle, c->slab));
  b.nx = c->g.nx; b.ny = c->g.ny; b.hd = c->g.hd; b.pitch = c->g.pitch;
  for (int i = 0; i < B_COUNT; i++) b.buf_offset[i] = (long long)((char *)c->buf[i] - c->slab);
  b.mail_offset = (long long)((char *)c->mail - c->slab);
  b.rank = c->rank; b.device = c->device;
  b.pid = (long long)getpid(); b.slab_ptr = (unsigned long long)(uintptr_t)c->slab;
  memcpy(blob, &b, sizeof b);
  return TL_OK;
}

extern "C" int tl_comm_unique_id(void *id128) {
#ifdef TL_WITH_NCCL
  ncclUniqueId id;
  if (!nccl_api().ok || nccl_api().GetUniqueId(&id) != ncclSuccess) return TL_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id128, &id, 128);
  return TL_OK;
#else
  (void)id128;
  return TL_ERR_COMM;
#endif
}

extern "C" int tl_comm_connect(tl_ctx *c, const void *all_blobs, const void *id128) {
  if (TL_IS_MULTI(c)) return tl_fail(c, TL_ERR_STATE, "tl_comm_connect: a tl_create_multi context wires its own tiles");
  if (!c) return TL_ERR_ARG;
  if (c->nranks == 1) { c->comm_ready = true; return TL_OK; }
  if (!all_blobs) return tl_fail(c, TL_ERR_ARG, "tl_comm_connect: the tiles' blobs are required");
  CU(c, cudaSetDevice(c->device));
  const CommBlob *blobs = (const CommBlob *)all_blobs;
  // every tile maps every other tile's slab: the mailboxes are all-to-all, the fields are
  // only touched on the eight neighbours
  CommDev hd;
  memset(&hd, 0, sizeof hd);
  hd.nranks = c->nranks; hd.rank = c->rank; hd.deferred = c->xchg_deferred;
  const long long me = (long long)getpid();
  for (int r = 0; r < c->nranks; r++) {
    if (blobs[r].rank != r) return tl_fail(c, TL_ERR_ARG, "tl_comm_connect: blobs are not in rank order");
    if (r == c->rank) { hd.mail[r] = c->mail; continue; }
    if (blobs[r].pid == me) {
      // a tile of this very process (several tiles driven by threads, possibly on ONE GPU: the
      // single-GPU test harness): its allocation is addressable as it is
      if (blobs[r].device != c->device) {
        cudaError_t pe = cudaDeviceEnablePeerAccess(blobs[r].device, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (pe != cudaSuccess) return tl_fail(c, TL_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", blobs[r].device, cudaGetErrorString(pe));
      }
      c->rank_slab[r] = (void *)(uintptr_t)blobs[r].slab_ptr;
      c->rank_ipc[r] = false;
    } else {
      CU(c, cudaIpcOpenMemHandle(&c->rank_slab[r], blobs[r].handle, cudaIpcMemLazyEnablePeerAccess));
      c->rank_ipc[r] = true;
    }
    hd.mail[r] = (MailSlot *)((char *)c->rank_slab[r] + blobs[r].mail_offset);
  }
  // the table goes up on the context's OWN stream (a synchronous cudaMemcpy from pageable memory runs on the legacy
  // default stream, which the non-blocking solver stream is not ordered with, and may return before its DMA has landed)
  static_assert(sizeof(CommDev) <= 64 * sizeof(double), "CommDev must fit the pinned scratch");
  if (getenv("TL_COMM_TABLE_LEGACY")) {
    CU(c, cudaMemcpy(c->d_comm, &hd, sizeof hd, cudaMemcpyHostToDevice));
  } else {
    memcpy(c->h_scal, &hd, sizeof hd);
    CU(c, cudaMemcpyAsync(c->d_comm, c->h_scal, sizeof hd, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
  }
  for (int r = 0; r < c->nranks; r++) c->rank_blob[r] = blobs[r];
  for (int s = 0; s < 4; s++) {
    const int nr = c->nbr_rank[s];
    if (nr < 0) continue;
    c->peer_blob[s] = blobs[nr];
    c->peer_slab[s] = c->rank_slab[nr];
  }
  // NCCL serves the legacy mode (comm_fused = 0) only; without an id the context is fused-only
  if (id128) {
#ifdef TL_WITH_NCCL
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    if (!nccl_api().ok) return tl_fail(c, TL_ERR_COMM, "libnccl.so.2 could not be loaded");
    ncclResult_t r = nccl_api().CommInitRank(&c->nccl, c->nranks, id, c->rank);
    if (r != ncclSuccess) return tl_fail(c, TL_ERR_COMM, "ncclCommInitRank: %s", nccl_api().GetErrorString(r));
#else
    return tl_fail(c, TL_ERR_COMM, "library built without NCCL");
#endif
  }
  c->comm_ready = true;
  return TL_OK;
}

// All-tiles sum of n doubles through the peer-mapped mailboxes (tl_tile_exchange): one small
// kernel, ~2 us, no library call -- what the fused mode uses outside the solver kernels (init,
// summaries, rendezvous).  Everything earlier kernels of this stream wrote (peer stores
// included) is ordered before the mailbox stores by the fence.sys + bar.sync at the start.
__global__ void k_tile_allreduce(const CommDev *cd, SolveState *st, const double *src, double *dst, int n) {
  __shared__ double sm[32];
  (void)tl_entry_scalars(cd, st);   // split exchange: collect what a loop kernel posted before exchanging again
  __syncthreads();
  if (threadIdx.x == 0) st->pend_target = TL_T_NONE;   // consumed: whoever writes red_* next is not patched over
  if (st->comm_error) return;
  if (threadIdx.x == blockDim.x - 1) __threadfence_system();
  for (int q = 0; q < n; q++) {
    const double v = (threadIdx.x == 0) ? src[q] : 0.0;
    const double t = tl_tile_exchange(cd, st, v, sm);
    if (threadIdx.x == 0) dst[q] = t;
  }
}

// Split exchange: publishes the total of the exchange the last loop kernel posted (nobody collected it yet) into the
// SolveState.  Enqueued wherever the host, or a kernel that reads the SolveState directly, comes next.
__global__ void k_xchg_finalize(const CommDev *cd, SolveState *st) {
  (void)tl_entry_scalars(cd, st);
  __syncthreads();
  if (threadIdx.x == 0) st->pend_target = TL_T_NONE;   // consumed: the SolveState holds the totals; later direct writers are not patched over
}
static bool xchg_deferred(const tl_ctx *c) { return c->nranks > 1 && c->comm_fused && c->xchg_deferred; }
static int xchg_finalize(tl_ctx *c) {
  if (!xchg_deferred(c)) return TL_OK;
  k_xchg_finalize<<<1, 32, 0, c->stream>>>(c->d_comm, c->st);
  c->launches++;
  CHECK_LAUNCH(c);
  return TL_OK;
}

// sum over tiles of n doubles living in device memory (stream ordered); out of place when
// src != dst.  With one tile it degenerates to a copy (or nothing).
static int allreduce2(tl_ctx *c, const double *src, double *dst, int n);
static int allreduce(tl_ctx *c, double *dev, int n) { return allreduce2(c, dev, dev, n); }
static int allreduce2(tl_ctx *c, const double *src, double *dst, int n) {
  if (c->nranks == 1) {
    if (src != dst) CU(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return TL_OK;
  }
  if (!c->comm_ready) return tl_fail(c, TL_ERR_STATE, "tile context used before tl_comm_connect");
  if (c->comm_fused) {
    k_tile_allreduce<<<1, 32, 0, c->stream>>>(c->d_comm, c->st, src, dst, n);
    c->launches++;
    CHECK_LAUNCH(c);
    return TL_OK;
  }
#ifdef TL_WITH_NCCL
  if (!c->nccl) return tl_fail(c, TL_ERR_STATE, "comm_fused = 0 needs NCCL: tl_comm_connect was called without an NCCL id");
  ncclResult_t r = nccl_api().AllReduce(src, dst, n, ncclDouble, ncclSum, c->nccl, c->stream);
  if (r != ncclSuccess) return tl_fail(c, TL_ERR_COMM, "ncclAllReduce: %s", nccl_api().GetErrorString(r));
  return TL_OK;
#else
  return tl_fail(c, TL_ERR_COMM, "library built without NCCL");
#endif
}

static PeerFace peer_face(tl_ctx *c, int side, int bufidx) {
  PeerFace f{nullptr, 0, 0, 0};
  if (c->nbr_rank[side] < 0 || !c->peer_slab[side]) return f;
  const CommBlob &b = c->peer_blob[side];
  f.f0 = (const double *)((char *)c->peer_slab[side] + b.buf_offset[bufidx]);
  f.nx = b.nx; f.ny = b.ny; f.pitch = b.pitch;
  return f;
}

// halo targets of buffer `bufidx` on the tile-internal sides (fused multi-GPU mode)
static Push push_for(tl_ctx *c, int bufidx) {
  Push p;
  memset(&p, 0, sizeof p);
  if (c->nranks == 1 || !c->comm_fused) return p;
  for (int s = 0; s < 4; s++) {
    if (c->nbr_rank[s] < 0 || !c->peer_slab[s]) continue;
    const CommBlob &b = c->peer_blob[s];
    p.s[s].f0 = (double *)((char *)c->peer_slab[s] + b.buf_offset[bufidx]);
    p.s[s].pitch = b.pitch; p.s[s].nx = b.nx; p.s[s].ny = b.ny;
  }
  return p;
}
// depth-k halo targets of buffer `bufidx` on the eight surrounding tiles
static Push8 push8_for(tl_ctx *c, int bufidx) {
  Push8 p;
  memset(&p, 0, sizeof p);
  if (c->nranks == 1 || !c->comm_fused) return p;
  for (int dy = -1; dy <= 1; dy++)
    for (int dx = -1; dx <= 1; dx++) {
      const int tx = c->cx + dx, ty = c->cy + dy;
      if ((dx == 0 && dy == 0) || tx < 0 || tx >= c->px || ty < 0 || ty >= c->py) continue;
      const int r = tx + ty * c->px;
      if (!c->rank_slab[r]) continue;
      const CommBlob &b = c->rank_blob[r];
      PushSide &t = p.s[(dy + 1) * 3 + (dx + 1)];
      t.f0 = (double *)((char *)c->rank_slab[r] + b.buf_offset[bufidx]);
      t.pitch = b.pitch; t.nx = b.nx; t.ny = b.ny;
    }
  return p;
}
static const CommDev *comm_dev(tl_ctx *c) { return (c->nranks > 1 && c->comm_fused) ? c->d_comm : nullptr; }
static bool legacy_comm(tl_ctx *c) { return c->nranks > 1 && !c->comm_fused; }

static int buf_index(tl_ctx *c, int f) {
  if (f == TL_P) return c->p_cur ? B_P1 : TL_P;
  if (f == TL_U) return c->u_cur ? B_U1 : TL_U;
  if (f == TL_SD) return c->sd_cur ? B_SD1 : TL_SD;
  return f;
}

// tile-internal halos of buffer `bufidx`; the caller guarantees (through a preceding
// allreduce or tl_comm barrier) that the neighbours have finished writing it.
static int pull_halo(tl_ctx *c, int bufidx, int depth) {
  if (c->nranks == 1) return TL_OK;
  const int total = 2 * depth * (c->g.nx + c->g.ny);
  const int grid = std::min((total + 255) / 256, c->num_sms * 4);
  k_pull_halo<<<grid, 256, 0, c->stream>>>(c->g, depth, c->buf[bufidx], peer_face(c, 0, bufidx),
                                           peer_face(c, 1, bufidx), peer_face(c, 2, bufidx), peer_face(c, 3, bufidx));
  c->launches++;
  CHECK_LAUNCH(c);
  return TL_OK;
}

static int tile_barrier(tl_ctx *c);
// tile-internal halos of buffer `bufidx`, `depth` cells deep INCLUDING the corner blocks (two
// phases with a rendezvous in between; see k_pull_halo_wide).  The caller guarantees the
// neighbours' interiors are final; ends with a rendezvous.
static int pull_halo_wide(tl_ctx *c, const int *bufs, int nbufs, int depth) {
  if (c->nranks == 1) return TL_OK;
  for (int phase = 0; phase < 2; phase++) {
    const int total = 2 * depth * (phase == 0 ? c->g.ny : c->g.nx + 2 * depth);
    const int grid = std::max(1, std::min((total + 255) / 256, c->num_sms * 4));
    for (int q = 0; q < nbufs; q++) {
      const int b = bufs[q];
      k_pull_halo_wide<<<grid, 256, 0, c->stream>>>(c->g, depth, phase, c->buf[b], peer_face(c, 0, b), peer_face(c, 1, b),
                                                    peer_face(c, 2, b), peer_face(c, 3, b));
      c->launches++;
      CHECK_LAUNCH(c);
    }
    TRY(tile_barrier(c));
  }
  return TL_OK;
}

// a device-side rendezvous of all tiles (1-double allreduce); used where a halo pull is not
// already ordered by a dot-product allreduce.
static int tile_barrier(tl_ctx *c) {
  if (c->nranks == 1) return TL_OK;
  return allreduce2(c, &c->st->barrier_zero, &c->st->barrier_out, 1);
}

// ---------------------------------------------------------------------------------------
// field transfer
// ---------------------------------------------------------------------------------------
extern "C" int tl_set_field(tl_ctx *c, int field, const double *host, long ld) {
  if (TL_IS_MULTI(c)) return multi_set_field(c, field, host, ld);
  if (!c || !host || field < 0 || field >= TL_NUM_FIELDS) return tl_fail(c, TL_ERR_ARG, "tl_set_field: bad argument");
  const Geo &g = c->g;
  const int x = g.nx + 2 * g.hd, y = g.ny + 2 * g.hd;
  if (ld < x) return tl_fail(c, TL_ERR_ARG, "tl_set_field: ld < x");
  CU(c, cudaSetDevice(c->device));
  double *dst = field_ptr(c, field) - (long)g.hd * g.pitch - g.hd;
  CU(c, cudaMemcpy2DAsync(dst, (size_t)g.pitch * 8, host, (size_t)ld * 8, (size_t)x * 8, y, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_get_field(tl_ctx *c, int field, double *host, long ld) {
  if (TL_IS_MULTI(c)) return multi_get_field(c, field, host, ld);
  if (!c || !host || field < 0 || field >= TL_NUM_FIELDS) return tl_fail(c, TL_ERR_ARG, "tl_get_field: bad argument");
  const Geo &g = c->g;
  const int x = g.nx + 2 * g.hd, y = g.ny + 2 * g.hd;
  if (ld < x) return tl_fail(c, TL_ERR_ARG, "tl_get_field: ld < x");
  CU(c, cudaSetDevice(c->device));
  const double *src = field_ptr(c, field) - (long)g.hd * g.pitch - g.hd;
  CU(c, cudaMemcpy2DAsync(host, (size_t)ld * 8, src, (size_t)g.pitch * 8, (size_t)x * 8, y, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_copy_field(tl_ctx *c, int dst_field, int src_field) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_copy_field(t, dst_field, src_field); });
  if (!c || dst_field < 0 || dst_field >= TL_NUM_FIELDS || src_field < 0 || src_field >= TL_NUM_FIELDS)
    return tl_fail(c, TL_ERR_ARG, "tl_copy_field: bad field");
  CU(c, cudaSetDevice(c->device));
  const Geo &g = c->g;
  const size_t off = (size_t)g.hd * g.pitch + TL_XPAD;
  CU(c, cudaMemcpyAsync(field_ptr(c, dst_field) - off, field_ptr(c, src_field) - off,
                        c->rows * g.pitch * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

// setchunkstate! on the device (src/chunk.jl:122-151)
extern "C" int tl_paint_states(tl_ctx *c, int nstates, const tl_state *states, double xmin, double ymin, double dx,
                               double dy, int x0, int y0) {
  if (TL_IS_MULTI(c))
    return multi_all(c, [&](tl_ctx *t, int i) {
      int tx = 0, ty = 0;
      multi_tile_offset(c, i, &tx, &ty);
      return tl_paint_states(t, nstates, states, xmin, ymin, dx, dy, x0 + tx, y0 + ty);
    });
  if (!c || !states || nstates < 1) return tl_fail(c, TL_ERR_ARG, "tl_paint_states: bad argument");
  if (nstates > TL_MAX_STATES) return tl_fail(c, TL_ERR_ARG, "tl_paint_states: more than %d states", TL_MAX_STATES);
  CU(c, cudaSetDevice(c->device));
  PaintParams P;
  memset(&P, 0, sizeof P);
  P.n = nstates; P.xmin = xmin; P.ymin = ymin; P.dx = dx; P.dy = dy; P.x0 = x0; P.y0 = y0;
  for (int q = 0; q < nstates; q++) {
    if (states[q].geometry < TL_GEOM_RECTANGULAR || states[q].geometry > TL_GEOM_POINT)
      return tl_fail(c, TL_ERR_ARG, "tl_paint_states: state %d has an unknown geometry", q + 1);
    P.s[q].density = states[q].density; P.s[q].energy = states[q].energy;
    P.s[q].xmin = states[q].xmin; P.s[q].ymin = states[q].ymin; P.s[q].xmax = states[q].xmax; P.s[q].ymax = states[q].ymax;
    P.s[q].radius = states[q].radius; P.s[q].geometry = states[q].geometry;
  }
  c->u_cur = 0;
  k_paint_states<<<c->basic_grid, TL_BASIC_THREADS, 0, c->stream>>>(c->g, P, c->buf[TL_DENSITY], c->buf[TL_ENERGY0], c->buf[TL_U]);
  c->launches++;
  CHECK_LAUNCH(c);
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

// ---------------------------------------------------------------------------------------
// per-function kernels
// ---------------------------------------------------------------------------------------
#define LAUNCH_BASIC(c, kern, ...)                                                      \
  do { kern<<<(c)->basic_grid, TL_BASIC_THREADS, 0, (c)->stream>>>(__VA_ARGS__);        \
       (c)->launches++; CHECK_LAUNCH(c); } while (0)

// tiles: did a rendezvous of this context time out?  (call with the stream idle; the flag is sticky)
static int check_comm(tl_ctx *c, const char *where) {
  if (c->nranks == 1) return TL_OK;
  SolveState *h = &c->h_st[0];
  CU(c, cudaMemcpyAsync(h, c->st, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (h->comm_error) {
    // where this tile's packets of its last blocking exchange went, against where the table says they belong
    std::string tgt;
    for (int r = 0; r < c->nranks; r++) {
      const MailSlot *base = (r == c->rank) ? c->mail : (const MailSlot *)((char *)c->rank_slab[r] + c->rank_blob[r].mail_offset);
      const long long off = (long long)(h->dbg_dst[r] - (unsigned long long)(uintptr_t)base);
      char b[64];
      snprintf(b, sizeof b, "%s%d:%s%lld", r ? " " : "", r, (off >= 0 && off < (long long)(2 * TL_MAX_RANKS * sizeof(MailSlot))) ? "+" : "STRAY", off);
      tgt += b;
    }
    const int who = h->comm_error - 1;
    return tl_fail(c, TL_ERR_COMM, "%s: tile exchange timed out: tile %d did not hear from tile %d in exchange %llu (slot held exchange numbers "
                   "%u/%u; this tile's packet offsets in the destination mailboxes: %s)", where, c->rank, who, h->xseq,
                   (unsigned)(h->dbg_seen[who & 15] >> 32), (unsigned)h->dbg_seen[who & 15], tgt.c_str());
  }
  return TL_OK;
}

static int xchg_finalize(tl_ctx *c);
static int read_scalars(tl_ctx *c, const double *dev, int n, double *out) {
  TRY(xchg_finalize(c));   // split exchange: a loop kernel may have posted the value that is read here
  CU(c, cudaMemcpyAsync(c->h_scal, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) out[i] = c->h_scal[i];
  return check_comm(c, "all-tiles sum");
}

static int halo_update_buf(tl_ctx *c, int bufidx, int depth) {
  const int total = 2 * depth * (c->g.nx + c->g.ny);
  const int grid = std::min((total + 255) / 256, c->num_sms * 4);
  if (c->g.phys) {
    k_halo_reflect<<<grid, 256, 0, c->stream>>>(c->g, depth, c->buf[bufidx]);
    c->launches++;
    CHECK_LAUNCH(c);
  }
  return pull_halo(c, bufidx, depth);
}

extern "C" int tl_halo_update(tl_ctx *c, unsigned field_mask, int depth) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_halo_update(t, field_mask, depth); });
  if (!c || depth < 1 || depth > c->g.hd) return tl_fail(c, TL_ERR_ARG, "tl_halo_update: depth must be in 1..halo_depth");
  CU(c, cudaSetDevice(c->device));
  TRY(tile_barrier(c));  // neighbours' interiors must be final before they are pulled
  for (int f = 0; f < TL_NUM_FIELDS; f++)
    if (field_mask & (1u << f)) TRY(halo_update_buf(c, buf_index(c, f), depth));
  TRY(tile_barrier(c));  // nobody overwrites an interior a neighbour is still reading
  CU(c, cudaStreamSynchronize(c->stream));
  return check_comm(c, "tl_halo_update");
}

// CG.init! (CG.jl:47-79), asynchronous part: leaves rro in st->red_rr (all-reduced).
static int cg_init_async(tl_ctx *c, int coef, double rx, double ry) {
  if (coef != TL_CONDUCTIVITY && coef != TL_RECIP_CONDUCTIVITY)
    return tl_fail(c, TL_ERR_ARG, "Coefficient %d is not valid", coef);  // CG.jl:48-50
  c->p_cur = c->u_cur = c->sd_cur = 0;
  const Geo &g = c->g;
  LAUNCH_BASIC(c, k_init_fields, g, coef, c->buf[TL_ENERGY], c->buf[TL_DENSITY], c->buf[TL_U], c->buf[TL_P],
               c->buf[TL_R], c->buf[TL_W]);
  LAUNCH_BASIC(c, k_init_k, g, rx, ry, c->buf[TL_W], c->buf[TL_KX], c->buf[TL_KY]);
  LAUNCH_BASIC(c, k_init_wrp, g, c->buf[TL_U], c->buf[TL_KX], c->buf[TL_KY], c->buf[TL_W], c->buf[TL_R],
               c->buf[TL_P], c->partials, &c->st->counter, &c->st->red_rr);
  return allreduce(c, &c->st->red_rr, 1);
}

extern "C" int tl_cg_init(tl_ctx *c, int coef, double rx, double ry, double *rro) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, rro, [&](tl_ctx *t, double *o) { return tl_cg_init(t, coef, rx, ry, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  TRY(cg_init_async(c, coef, rx, ry));
  double v;
  TRY(read_scalars(c, &c->st->red_rr, 1, &v));
  if (rro) *rro = v;
  return TL_OK;
}

extern "C" int tl_cg_calc_w(tl_ctx *c, double *pw) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, pw, [&](tl_ctx *t, double *o) { return tl_cg_calc_w(t, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_calc_w, c->g, field_ptr(c, TL_P), c->buf[TL_KX], c->buf[TL_KY], c->buf[TL_W], c->partials,
               &c->st->counter, &c->st->red_pw);
  TRY(allreduce(c, &c->st->red_pw, 1));
  double v;
  TRY(read_scalars(c, &c->st->red_pw, 1, &v));
  if (pw) *pw = v;
  return TL_OK;
}

extern "C" int tl_cg_calc_ur(tl_ctx *c, double alpha, double *rrn) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, rrn, [&](tl_ctx *t, double *o) { return tl_cg_calc_ur(t, alpha, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_calc_ur, c->g, alpha, field_ptr(c, TL_P), c->buf[TL_W], field_ptr(c, TL_U), c->buf[TL_R],
               c->partials, &c->st->counter, &c->st->red_aux[0]);
  TRY(allreduce(c, &c->st->red_aux[0], 1));
  double v;
  TRY(read_scalars(c, &c->st->red_aux[0], 1, &v));
  if (rrn) *rrn = v;
  return TL_OK;
}

extern "C" int tl_cg_calc_p(tl_ctx *c, double beta) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_cg_calc_p(t, beta); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_calc_p, c->g, beta, c->buf[TL_R], field_ptr(c, TL_P));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_copy_u(tl_ctx *c) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_copy_u(t); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_copy, c->g, 0, field_ptr(c, TL_U), c->buf[TL_U0]);
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

static int residual_async(tl_ctx *c) {
  LAUNCH_BASIC(c, k_residual, c->g, field_ptr(c, TL_U), c->buf[TL_U0], c->buf[TL_KX], c->buf[TL_KY], c->buf[TL_R]);
  return TL_OK;
}
extern "C" int tl_calc_residual(tl_ctx *c) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_calc_residual(t); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  TRY(residual_async(c));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_finalise(tl_ctx *c) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_finalise(t); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_finalise, c->g, field_ptr(c, TL_U), c->buf[TL_DENSITY], c->buf[TL_ENERGY]);
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_solve_finished(tl_ctx *c, int check_result) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_solve_finished(t, check_result); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  if (check_result) TRY(residual_async(c));
  LAUNCH_BASIC(c, k_finalise, c->g, field_ptr(c, TL_U), c->buf[TL_DENSITY], c->buf[TL_ENERGY]);
  CU(c, cudaStreamSynchronize(c->stream));
  return tl_halo_update(c, TL_MASK(TL_ENERGY), 1);
}

static int norm2_async(tl_ctx *c, int field, double *dev_out) {
  LAUNCH_BASIC(c, k_norm2, c->g, field_ptr(c, field), c->partials, &c->st->counter, dev_out);
  return allreduce(c, dev_out, 1);
}
extern "C" int tl_norm2(tl_ctx *c, int field, double *out) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, out, [&](tl_ctx *t, double *o) { return tl_norm2(t, field, o); });
  if (!c || field < 0 || field >= TL_NUM_FIELDS) return tl_fail(c, TL_ERR_ARG, "tl_norm2: bad field");
  CU(c, cudaSetDevice(c->device));
  TRY(norm2_async(c, field, &c->st->red_aux[0]));
  double v;
  TRY(read_scalars(c, &c->st->red_aux[0], 1, &v));
  if (out) *out = v;
  return TL_OK;
}

extern "C" int tl_cheby_init(tl_ctx *c, double theta, double *bb) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, bb, [&](tl_ctx *t, double *o) { return tl_cheby_init(t, theta, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  TRY(norm2_async(c, TL_U0, &c->st->red_aux[0]));   // Cheby.jl:68
  LAUNCH_BASIC(c, k_cheby_wrp, c->g, 1, theta, 0.0, 0.0, field_ptr(c, TL_U), c->buf[TL_U0], c->buf[TL_KX],
               c->buf[TL_KY], c->buf[TL_W], c->buf[TL_R], field_ptr(c, TL_P));
  LAUNCH_BASIC(c, k_add, c->g, field_ptr(c, TL_P), field_ptr(c, TL_U));
  double v;
  TRY(read_scalars(c, &c->st->red_aux[0], 1, &v));
  if (bb) *bb = v;
  return tl_halo_update(c, TL_MASK(TL_U), 1);        // Cheby.jl:78
}

extern "C" int tl_cheby_iterate(tl_ctx *c, double alpha, double beta, int calc_2norm, double *error) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, error, [&](tl_ctx *t, double *o) { return tl_cheby_iterate(t, alpha, beta, calc_2norm, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_cheby_wrp, c->g, 0, 1.0, alpha, beta, field_ptr(c, TL_U), c->buf[TL_U0], c->buf[TL_KX],
               c->buf[TL_KY], c->buf[TL_W], c->buf[TL_R], field_ptr(c, TL_P));
  LAUNCH_BASIC(c, k_add, c->g, field_ptr(c, TL_P), field_ptr(c, TL_U));
  if (calc_2norm) {
    TRY(norm2_async(c, TL_R, &c->st->red_aux[0]));
    double v;
    TRY(read_scalars(c, &c->st->red_aux[0], 1, &v));
    if (error) *error = v;
  } else {
    CU(c, cudaStreamSynchronize(c->stream));
  }
  return TL_OK;
}

extern "C" int tl_ppcg_init_sd(tl_ctx *c, double theta) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_ppcg_init_sd(t, theta); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_ppcg_init_sd, c->g, theta, c->buf[TL_R], field_ptr(c, TL_SD));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_ppcg_inner(tl_ctx *c, const double *alphas, const double *betas, int nsteps) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_ppcg_inner(t, alphas, betas, nsteps); });
  if (!c || !alphas || !betas || nsteps < 0) return tl_fail(c, TL_ERR_ARG, "tl_ppcg_inner: bad argument");
  CU(c, cudaSetDevice(c->device));
  for (int pp = 0; pp < nsteps; pp++) {
    TRY(tl_halo_update(c, TL_MASK(TL_SD), 1));   // PPCG.jl:76
    LAUNCH_BASIC(c, k_ppcg_inner1, c->g, field_ptr(c, TL_SD), c->buf[TL_KX], c->buf[TL_KY], c->buf[TL_R], field_ptr(c, TL_U));
    LAUNCH_BASIC(c, k_ppcg_inner2, c->g, alphas[pp], betas[pp], c->buf[TL_R], field_ptr(c, TL_SD));
  }
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_field_summary(tl_ctx *c, double cell_volume, double *vol, double *mass, double *ie, double *temp) {
  if (TL_IS_MULTI(c)) return multi_field_summary(c, cell_volume, vol, mass, ie, temp);
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_field_summary, c->g, cell_volume, c->buf[TL_DENSITY], c->buf[TL_ENERGY0], field_ptr(c, TL_U),
               c->partials, &c->st->counter, c->st->red_aux);
  TRY(allreduce(c, c->st->red_aux, 4));
  double v[4];
  TRY(read_scalars(c, c->st->red_aux, 4, v));
  if (vol) *vol = v[0];
  if (mass) *mass = v[1];
  if (ie) *ie = v[2];
  if (temp) *temp = v[3];
  return TL_OK;
}

// ---------------------------------------------------------------------------------------
// whole-solve drivers
// ---------------------------------------------------------------------------------------
__global__ void k_state_begin(SolveState *st, StopCfg cfg, int iter, double theta, int inner_steps) {
  st->cfg = cfg;
  st->iter = iter;
  st->counter = 0u;
  st->theta = theta;
  st->inner_steps = inner_steps;
  st->inner_pp = 0;
}
__global__ void k_state_cheby(SolveState *st, double theta, double eps, int tt0, int max_tt) {
  st->theta = theta;
  st->eps_cheby = eps;
  st->cheby_step = 0;
  st->cheby_pairs = 0;
  st->cheby_est = INT_MAX;
  st->cheby_tt0 = tt0;
  st->cheby_max_tt = max_tt;
  st->counter = 0u;
}
__global__ void k_state_set_est(SolveState *st, int est) { st->cheby_est = est; }
__global__ void k_state_set_step(SolveState *st, int step) { st->cheby_step = step; st->cheby_pairs = 0; }

static CgAParams cg_a_params(tl_ctx *c) {
  CgAParams P;
  P.g = c->g; P.t = c->tiling; P.st = c->st; P.hist_rr = c->hist_rr; P.hist_pw = c->hist_pw;
  P.r = c->buf[TL_R]; P.p0 = c->buf[TL_P]; P.p1 = c->buf[B_P1]; P.u = c->buf[TL_U];
  P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.w = c->buf[TL_W]; P.partials = c->partials;
  P.single = c->nranks == 1;
  P.hint_keep = c->hint_keep; P.hint_stream = c->hint_stream;
  P.cd = comm_dev(c); P.push_p0 = push_for(c, TL_P); P.push_p1 = push_for(c, B_P1);
  P.lazy_role = 0;
  return P;
}
static CgBParams cg_b_params(tl_ctx *c) {
  CgBParams P;
  P.g = c->g; P.t = c->pw_tiling; P.st = c->st; P.hist_pw = c->hist_pw;
  P.r = c->buf[TL_R]; P.w = c->buf[TL_W]; P.partials = c->partials;
  P.single = c->nranks == 1;
  P.hint_keep = c->hint_keep; P.hint_stream = c->hint_stream; P.reverse = c->b_reverse;
  P.cd = comm_dev(c); P.push_r = push_for(c, TL_R);
  return P;
}

// kernel A in the configured flavour (register double-buffering or cp.async ring)
template <int U, int S, int MINB>
static int launch_ring(tl_ctx *c, const CgAParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_CGA_STAGE_BYTES(U);
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cg_fused_w_ring<U, S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_cg_fused_w_ring<U, S, MINB>, c->fused_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
// ---- TMA flavour of kernel A (option a_tma): one 2-D tensor map per buffer, encoded once per context ----
typedef CUresult (*tl_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int tma_prepare(tl_ctx *c) {
  if (c->tma_ready) return TL_OK;
  static tl_encode_tiled_fn encode = nullptr;
  if (!encode) {   // the driver entry point, without linking libcuda
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CU(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return tl_fail(c, TL_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = (tl_encode_tiled_fn)fn;
  }
  const int bufs[TMA_NMAPS] = {TL_R, TL_P, B_P1, TL_KY, TL_KX, TL_U};
  for (int q = 0; q < TMA_NMAPS; q++) {
    // the whole padded buffer: dim0 = the padded row (pitch doubles), dim1 = rows incl. halos; box = 68 x 1
    void *base = (double *)c->slab + (size_t)bufs[q] * c->buf_doubles;
    const cuuint64_t dims[2] = {(cuuint64_t)c->g.pitch, (cuuint64_t)c->rows};
    const cuuint64_t strides[1] = {(cuuint64_t)c->g.pitch * sizeof(double)};
    const cuuint32_t box[2] = {TL_TMA_BOX, 1}, estr[2] = {1, 1};
    const CUresult r = encode(&c->tma_maps[q], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return tl_fail(c, TL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for buffer %d", (int)r, bufs[q]);
  }
  c->tma_ready = true;
  return TL_OK;
}
template <bool U, int S, int MINB>
static int launch_cg_a_tma(tl_ctx *c, const CgAParams &A) {
  TRY(tma_prepare(c));
  CgATmaParams P;
  P.a = A;
  memcpy(P.maps, c->tma_maps, sizeof P.maps);
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_TMA_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cg_fused_w_tma<U, S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_cg_fused_w_tma<U, S, MINB>, c->fused_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}

template <int U>
static int launch_cg_a(tl_ctx *c, int lazy_role = 0) {
  CgAParams P = cg_a_params(c);
  P.lazy_role = lazy_role;
  if (c->a_tma) {
    if constexpr (U == TL_U_LAZY) {
      return tl_fail(c, TL_ERR_STATE, "internal: the TMA flavour has no lazy-u mode");
    } else {
      if (c->a_tma == 3) TRY((launch_cg_a_tma<U != TL_U_NONE, 3, 2>(c, P)));
      else TRY((launch_cg_a_tma<U != TL_U_NONE, 4, 2>(c, P)));
      CHECK_LAUNCH(c);
      return TL_OK;
    }
  }
  switch (c->ring_eff) {
    case 3:
      if constexpr (U == TL_U_LAZY) {
        if (c->lazy_heavy_ctas == 2) { TRY((launch_ring<U, 3, 2>(c, P))); break; }
      }
      TRY((launch_ring<U, 3, 3>(c, P)));
      break;
    case 4: TRY((launch_ring<U, 4, 2>(c, P))); break;
    case 6: TRY((launch_ring<U, 6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  CHECK_LAUNCH(c);
  return TL_OK;
}

// resident CTAs per SM of the CG loop's kernel A as configured (read-back "cg_a_blocks_per_sm")
template <int U, int S, int MINB>
static int ring_occupancy(tl_ctx *c, int *out) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_CGA_STAGE_BYTES(U);
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cg_fused_w_ring<U, S, MINB>, smem, &prepared));
  CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, k_cg_fused_w_ring<U, S, MINB>, TL_FUSED_THREADS, smem));
  return TL_OK;
}
static int cg_a_occupancy(tl_ctx *c, int *out) {
  const bool lazy = cg_u_mode(c) == TL_U_LAZY;
  switch (c->ring_eff) {
    case 3:
      if (lazy) return c->lazy_heavy_ctas == 2 ? ring_occupancy<TL_U_LAZY, 3, 2>(c, out) : ring_occupancy<TL_U_LAZY, 3, 3>(c, out);
      return ring_occupancy<TL_U_EVERY, 3, 3>(c, out);
    case 4: return lazy ? ring_occupancy<TL_U_LAZY, 4, 2>(c, out) : ring_occupancy<TL_U_EVERY, 4, 2>(c, out);
    case 6: return lazy ? ring_occupancy<TL_U_LAZY, 6, 1>(c, out) : ring_occupancy<TL_U_EVERY, 6, 1>(c, out);
  }
  return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
}

template <bool FIRST, int S, int MINB>
static int launch_cheby_ring(tl_ctx *c, const ChebyParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cheby_fused_ring<FIRST, S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_cheby_fused_ring<FIRST, S, MINB>, c->fused_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static ChebyParams cheby_params(tl_ctx *c);
template <bool FIRST>
static int launch_cheby(tl_ctx *c) {
  const ChebyParams P = cheby_params(c);
  switch (c->ring_eff) {
    case 3: TRY((launch_cheby_ring<FIRST, 3, 3>(c, P))); break;
    case 4: TRY((launch_cheby_ring<FIRST, 4, 2>(c, P))); break;
    case 6: TRY((launch_cheby_ring<FIRST, 6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  CHECK_LAUNCH(c);
  return TL_OK;
}
template <int S, int MINB>
static int launch_ppcg_inner_ring(tl_ctx *c, const PpcgInnerParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_ppcg_inner_ring<S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_ppcg_inner_ring<S, MINB>, c->fused_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static PpcgInnerParams ppcg_inner_params(tl_ctx *c);
static int launch_ppcg_inner(tl_ctx *c) {
  const PpcgInnerParams P = ppcg_inner_params(c);
  switch (c->ring_eff) {
    case 3: TRY((launch_ppcg_inner_ring<3, 3>(c, P))); break;
    case 4: TRY((launch_ppcg_inner_ring<4, 2>(c, P))); break;
    case 6: TRY((launch_ppcg_inner_ring<6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  CHECK_LAUNCH(c);
  return TL_OK;
}

// kernel B in the configured flavour
template <int D, int MINB>
static int launch_b_ring(tl_ctx *c, const CgBParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * D * TL_BRING_ROW_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cg_fused_r_ring<D, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_cg_fused_r_ring<D, MINB>, c->pw_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static int launch_cg_b(tl_ctx *c) {
  const CgBParams P = cg_b_params(c);
  switch (c->b_ring) {
    case 0: {
      static std::atomic<unsigned long long> prepared{0};
      TRY(tl_prepare_smem(c, k_cg_fused_r, 0, &prepared));
      CU(c, tl_launch(c, k_cg_fused_r, c->pw_grid, TL_FUSED_THREADS, 0, P));
      break;
    }
    case 6: TRY((launch_b_ring<6, 4>(c, P))); break;
    case 8: TRY((launch_b_ring<8, 3>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: b_ring %d", c->b_ring);
  }
  CHECK_LAUNCH(c);
  return TL_OK;
}

// One CG iteration on the stream: kernel A, kernel B.  Tiled (fused mode): the same two kernels --
// they push their edge cells into the neighbours' halos and sum pw / rr over the tiles in their
// tails.  Legacy mode (comm_fused = 0, kept for A/B measurements): halo-pull kernels and NCCL.
static int enqueue_cg_iteration(tl_ctx *c) {
  const bool legacy = legacy_comm(c);
  if (legacy) {
    // r is final everywhere after the previous allreduce(rr); p_old after the one before.
    TRY(pull_halo(c, TL_R, 1));
    TRY(pull_halo(c, TL_P, 1));
    TRY(pull_halo(c, B_P1, 1));
  }
  if (cg_u_mode(c) == TL_U_LAZY) {   // even position in the chunk = even iteration of the phase: both pending u updates
    if ((c->cg_seq++ & 1) == 0) TRY(launch_cg_a<TL_U_LAZY>(c, 2));
    else TRY(launch_cg_a<TL_U_NONE>(c, 1));
  } else {
    TRY(launch_cg_a<TL_U_EVERY>(c));
  }
  if (legacy) TRY(allreduce2(c, &c->st->red_pw_local, &c->st->red_pw, 1));
  TRY(launch_cg_b(c));
  if (legacy) TRY(allreduce2(c, &c->st->red_rr_local, &c->st->red_rr, 1));
  c->launches += 2;
  return TL_OK;
}

template <typename F>
static int build_graph(tl_ctx *c, cudaGraphExec_t *exec, int reps, F enqueue_one) {
  cudaGraph_t graph = nullptr;
  CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  int rc = TL_OK;
  const long long saved = c->launches;
  c->cg_seq = 0;
  for (int i = 0; i < reps && rc == TL_OK; i++) rc = enqueue_one();
  if (rc == TL_OK) rc = xchg_finalize(c);   // the host reads the state after every chunk
  c->launches = saved;
  cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) return tl_fail(c, TL_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return tl_fail(c, TL_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  return TL_OK;
}

// Runs chunks of `chunk_iters` iterations ahead of the host; `stopped(state)` is the host's
// copy of the device-side stop rule.  On return the stream is idle and *final holds the state.
template <typename Enq, typename Stop>
static int run_chunks(tl_ctx *c, cudaGraphExec_t *exec, int *exec_iters, int chunk_iters, long long launches_per_iter,
                      Enq enqueue_one, Stop stopped, SolveState *final_state) {
  if (c->use_graph && (!*exec || *exec_iters != chunk_iters)) {
    if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
    TRY(build_graph(c, exec, chunk_iters, enqueue_one));
    *exec_iters = chunk_iters;
  }
  int k = 0;
  for (;; k++) {
    if (c->use_graph) {
      CU(c, cudaGraphLaunch(*exec, c->stream));
      c->launches += launches_per_iter * chunk_iters + (xchg_deferred(c) ? 1 : 0);
    } else {
      c->cg_seq = 0;
      for (int i = 0; i < chunk_iters; i++) TRY(enqueue_one());
      TRY(xchg_finalize(c));
    }
    CU(c, cudaMemcpyAsync(&c->h_st[k & 1], c->st, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev[k & 1], c->stream));
    if (k >= 1) {
      CU(c, cudaEventSynchronize(c->ev[(k - 1) & 1]));
      if (stopped(c->h_st[(k - 1) & 1]) || c->h_st[(k - 1) & 1].comm_error || c->h_st[(k - 1) & 1].sched_error) break;
    }
  }
  CU(c, cudaStreamSynchronize(c->stream));
  *final_state = c->h_st[k & 1];
  if (final_state->sched_error) return tl_fail(c, TL_ERR_STATE, "internal: a kernel of the lazy-u CG loop ran at an iteration of the wrong parity");
  if (final_state->comm_error)
    return tl_fail(c, TL_ERR_COMM, "tile exchange timed out: tile %d did not hear from tile %d in exchange %llu (a neighbour tile did not "
                   "reach the same kernel); state: iter %d, cheby_step %d, cheby_pairs %d, inner_pp %d", c->rank,
                   final_state->comm_error - 1, final_state->xseq, final_state->iter, final_state->cheby_step,
                   final_state->cheby_pairs, final_state->inner_pp);
  if (!stopped(*final_state)) return tl_fail(c, TL_ERR_STATE, "internal: chunk loop ended before the stop rule fired");
  return TL_OK;
}

// CG preamble shared by the three solvers: CG.init!, haloupdate!([:u,:p]), copyu!  (CG.jl:9-12;
// SURVEY Appendix A #12 for Cheby/PPCG)
static int solve_preamble(tl_ctx *c, int coef, double rx, double ry, const StopCfg &cfg) {
  TRY(cg_init_async(c, coef, rx, ry));
  k_state_begin<<<1, 1, 0, c->stream>>>(c->st, cfg, 0, 0.0, 0);
  c->launches++;
  CHECK_LAUNCH(c);
  if (c->nranks > 1) TRY(tile_barrier(c));
  TRY(halo_update_buf(c, TL_U, 1));
  TRY(halo_update_buf(c, TL_P, 1));
  LAUNCH_BASIC(c, k_copy, c->g, 0, c->buf[TL_U], c->buf[TL_U0]);
  return TL_OK;
}

// The CG loop as one persistent cooperative kernel (single tile; option cg_persist).
template <int S, int MINB>
static int launch_cg_persist(tl_ctx *c, const CgPersistParams &P0) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cg_persist<S, MINB>, smem, &prepared));
  int per_sm = 0;
  CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_persist<S, MINB>, TL_FUSED_THREADS, smem));
  if (per_sm < 1) return tl_fail(c, TL_ERR_STATE, "persistent CG kernel does not fit on an SM");
  CgPersistParams P = P0;
  const int grid = std::max(1, std::min(per_sm * c->num_sms, std::max(P.nitems_a, P.nitems_b)));
  CU(c, cudaMemsetAsync(c->psync, 0, sizeof(PersistSync), c->stream));
  void *args[] = {(void *)&P};
  CU(c, cudaLaunchCooperativeKernel((const void *)k_cg_persist<S, MINB>, dim3((unsigned)grid), dim3(TL_FUSED_THREADS), args,
                                    (size_t)smem, c->stream));
  return TL_OK;
}

static int cg_phase_persist(tl_ctx *c, SolveState *fin) {
  CgPersistParams P;
  P.A = cg_a_params(c);
  P.B = cg_b_params(c);
  P.sync = c->psync;
  P.part_a = c->partials;
  P.part_b = c->partials + TL_MAX_GRID;
  P.nitems_a = c->fused_grid;
  P.nitems_b = c->pw_grid;
  switch (c->ring_eff) {
    case 3: TRY((launch_cg_persist<3, 3>(c, P))); break;
    case 4: TRY((launch_cg_persist<4, 2>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "cg_persist needs ring depth 3 or 4 (is %d)", c->ring_eff);
  }
  c->launches += 1;
  int aborted = 0;
  CU(c, cudaMemcpyAsync(&c->h_st[0], c->st, sizeof(SolveState), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(&aborted, &c->psync->abort, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  *fin = c->h_st[0];
  if (aborted) return tl_fail(c, TL_ERR_STATE, "persistent CG kernel: grid barrier timed out");
  if (!tl_should_stop(fin->iter, fin->red_rr, fin->cfg))
    return tl_fail(c, TL_ERR_STATE, "internal: persistent CG kernel ended before the stop rule fired");
  return TL_OK;
}

static int cg_phase(tl_ctx *c, SolveState *fin) {
  if (c->cg_persist && c->nranks == 1) return cg_phase_persist(c, fin);
  auto enq = [&]() { return enqueue_cg_iteration(c); };
  auto stop = [&](const SolveState &s) { return tl_should_stop(s.iter, s.red_rr, s.cfg); };
  return run_chunks(c, &c->g_cg, &c->g_cg_iters, c->graph_iters, 2, enq, stop, fin);
}

// flush the deferred p/u update of the last iteration and make buffer 0 the current p
static int cg_flush(tl_ctx *c, int iters_done, bool update_u) {
  CgAParams P = cg_a_params(c);
  P.t = c->pw_tiling;
  if (!update_u) k_cg_flush<TL_U_NONE><<<c->pw_grid, TL_FUSED_THREADS, 0, c->stream>>>(P);
  else if (cg_u_mode(c) == TL_U_LAZY) k_cg_flush<TL_U_LAZY><<<c->pw_grid, TL_FUSED_THREADS, 0, c->stream>>>(P);
  else k_cg_flush<TL_U_EVERY><<<c->pw_grid, TL_FUSED_THREADS, 0, c->stream>>>(P);
  c->launches++;
  CHECK_LAUNCH(c);
  c->p_cur = iters_done & 1;
  if (c->nranks > 1) {
    // haloupdate!(.., [:u,:p]) of the last iteration on the tile-internal sides
    TRY(tile_barrier(c));
    if (update_u) TRY(pull_halo(c, TL_U, 1));
    TRY(pull_halo(c, c->p_cur ? B_P1 : TL_P, 1));
    TRY(tile_barrier(c));
  }
  if (c->p_cur) {  // keep pointers canonical for the graphs of the next phase / solve
    LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[B_P1], c->buf[TL_P]);
    c->p_cur = 0;
  }
  return TL_OK;
}

static int fetch_cg_coefficients(tl_ctx *c, int iters, std::vector<double> &al, std::vector<double> &be) {
  std::vector<double> rr(iters + 1), pw(iters + 1);
  CU(c, cudaMemcpyAsync(rr.data(), c->hist_rr, (iters + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(pw.data(), c->hist_pw, (iters + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  al.resize(iters); be.resize(iters);
  for (int t = 1; t <= iters; t++) {
    al[t - 1] = rr[t - 1] / pw[t];   // CG.jl:35
    be[t - 1] = rr[t] / rr[t - 1];   // CG.jl:39
  }
  return TL_OK;
}

static void finish_timing(tl_ctx *c, tl_solve_info *info, long long launches0) {
  cudaEventRecord(c->ev_stop, c->stream);
  cudaEventSynchronize(c->ev_stop);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop);
  info->solve_ms = ms;
  c->last_cg_ms = ms;
  if (c->phase_marked) {
    float pms = 0.f;
    if (cudaEventElapsedTime(&pms, c->ev_start, c->ev_phase) == cudaSuccess) c->last_cg_ms = pms;
    c->phase_marked = false;
  }
  info->kernel_launches = c->launches - launches0;
}

extern "C" int tl_cg_solve(tl_ctx *c, int coef, double rx, double ry, double eps, int max_iters,
                           tl_solve_info *info, double *cg_alphas, double *cg_betas) {
  if (TL_IS_MULTI(c))
    return multi_solve(c, info, [&](tl_ctx *t, tl_solve_info *o, int i) {
      return tl_cg_solve(t, coef, rx, ry, eps, max_iters, o, i == 0 ? cg_alphas : nullptr, i == 0 ? cg_betas : nullptr);
    });
  if (!c || !info) return TL_ERR_ARG;
  memset(info, 0, sizeof *info);
  CU(c, cudaSetDevice(c->device));
  max_iters = std::min(max_iters, c->max_iters);
  const long long l0 = c->launches;
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  StopCfg cfg{max_iters, TL_CONV_SQRT, INT_MAX, 0, eps, 0.0};
  TRY(solve_preamble(c, coef, rx, ry, cfg));
  SolveState fin;
  TRY(cg_phase(c, &fin));
  TRY(cg_flush(c, fin.iter, true));
  // hist_rr[iter] is written by the flush kernel
  std::vector<double> al, be;
  TRY(fetch_cg_coefficients(c, fin.iter, al, be));
  finish_timing(c, info, l0);
  if (cg_alphas) memcpy(cg_alphas, al.data(), al.size() * sizeof(double));
  if (cg_betas) memcpy(cg_betas, be.data(), be.size() * sizeof(double));
  info->iters = info->cg_iters = fin.iter;
  info->error = fin.iter > 0 ? fin.red_rr : TL_ERROR_START;
  return TL_OK;
}

static ChebyParams cheby_params(tl_ctx *c) {
  ChebyParams P;
  P.g = c->g; P.t = c->tiling; P.st = c->st; P.alphas = c->ch_alphas; P.betas = c->ch_betas;
  P.u0 = c->buf[TL_U0]; P.ua = c->buf[TL_U]; P.ub = c->buf[B_U1]; P.p = c->buf[TL_P];
  P.w = c->buf[TL_W]; P.r = c->buf[TL_R]; P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.partials = c->partials;
  P.single = c->nranks == 1;
  P.cd = comm_dev(c); P.push_ua = push_for(c, TL_U); P.push_ub = push_for(c, B_U1);
  P.p1 = c->buf[B_P1];
  return P;
}

// One Chebyshev iteration = one kernel.  Tiled (fused mode): the kernel pushes the edge cells of
// u' into the neighbours' halos; its tail exchange is the norm sum on norm iterations and a plain
// completion barrier otherwise.  Legacy mode: halo pulls + an NCCL allreduce as the rendezvous.
static int enqueue_cheby_iteration(tl_ctx *c) {
  const bool legacy = legacy_comm(c);
  if (legacy) {
    TRY(pull_halo(c, TL_U, 1));
    TRY(pull_halo(c, B_U1, 1));
  }
  TRY(launch_cheby<false>(c));
  if (legacy) TRY(allreduce2(c, &c->st->red_norm_local, &c->st->red_norm, 1));
  c->launches++;
  return TL_OK;
}

// Can the pair kernels run on this context's tiles?  (depth-2 halos, every tile at least 2 x 2 cells, fused exchange)
static bool pairs_on_tiles(const tl_ctx *c) {
  if (!(c->pair_tiled && c->nranks > 1 && c->comm_fused && c->g.hd >= 2)) return false;
  for (int r = 0; r < c->nranks; r++)
    if (c->rank_blob[r].nx < 2 || c->rank_blob[r].ny < 2) return false;
  return true;
}
// Two Chebyshev iterations = one kernel (option cheby_pair; k_cheby_pair_ring).  On tiles (TILED) the kernel
// pushes u' two cells and p' one cell deep into the eight surrounding tiles and its tail is the rendezvous.
static bool cheby_pairs_tiled(const tl_ctx *c) { return c->cheby_pair && pairs_on_tiles(c); }
template <int S, int MINB, bool TILED>
static int launch_cheby_pair_ring(tl_ctx *c, const ChebyPairParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_cheby_pair_ring<S, MINB, TILED>, smem, &prepared));
  CU(c, tl_launch(c, k_cheby_pair_ring<S, MINB, TILED>, c->pair_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static int enqueue_cheby_pair(tl_ctx *c) {
  ChebyPairParams P;
  P.c = cheby_params(c);
  P.c.t = c->pair_tiling;
  if (c->nranks > 1) {
    P.push_ua = push8_for(c, TL_U); P.push_ub = push8_for(c, B_U1);
    P.push_p0 = push8_for(c, TL_P); P.push_p1 = push8_for(c, B_P1);
    if (c->pair_stages == 5) TRY((launch_cheby_pair_ring<5, 2, true>(c, P)));
    else TRY((launch_cheby_pair_ring<4, 2, true>(c, P)));
  } else {
    memset(&P.push_ua, 0, 4 * sizeof(Push8));
    if (c->pair_stages == 5) TRY((launch_cheby_pair_ring<5, 2, false>(c, P)));
    else TRY((launch_cheby_pair_ring<4, 2, false>(c, P)));
  }
  CHECK_LAUNCH(c);
  c->launches++;
  return TL_OK;
}
// Halos the first pair kernel of a phase reads on tile-internal sides: the given buffers two cells deep, corner
// blocks included, plus the neighbours' PHYSICAL-top halo row in the tile-internal halo columns (ky(.., ny) is read
// by the redundant cells next to a physical top; tests/emulation/emulate_pair_tiled.py).  Ends with a rendezvous.
static int pair_tiled_fill_halos(tl_ctx *c, const int *bufs, int nbufs) {
  TRY(tile_barrier(c));
  TRY(pull_halo_wide(c, bufs, nbufs, 2));
  for (int q = 0; q < nbufs; q++) {
    k_pull_halo_cols_top<<<1, 32, 0, c->stream>>>(c->g, 2, c->buf[bufs[q]], peer_face(c, 0, bufs[q]), peer_face(c, 1, bufs[q]));
    c->launches++;
    CHECK_LAUNCH(c);
  }
  TRY(tile_barrier(c));
  return TL_OK;
}

// may the next kernel be a pair?  (the host's copy of the kernel's entry conditions)
static bool cheby_pair_allowed(const SolveState &s) {
  if (tl_cheby_should_stop(s)) return false;
  const int ttA = s.cheby_tt0 + s.cheby_step - 1;
  if (ttA + 1 > s.cheby_max_tt) return false;
  return !tl_cheby_is_norm_iter(s.cheby_step, s.cheby_tt0, s.cheby_est);
}

// common switch bookkeeping of Cheby.solve!/PPCG.solve! (Cheby.jl:25-29, PPCG.jl:25-30)
static StopCfg switch_cfg(int max_iters, double eps, int presteps, double epslim, int errorswitch) {
  StopCfg cfg;
  cfg.max_iters = max_iters;
  cfg.conv_mode = TL_CONV_ABS;
  cfg.first_it = 0;
  cfg.eps = eps;
  cfg.switch_min = errorswitch ? TL_CGEIGENITERS : std::max(presteps, 0);
  cfg.switch_thresh = errorswitch ? epslim : TL_ERROR_SWITCH_MAX;
  if (cfg.switch_min < 1) cfg.switch_min = 1;   // `error` is ERROR_START until one CG iteration ran
  return cfg;
}

extern "C" int tl_cheby_solve(tl_ctx *c, int coef, double rx, double ry, double eps, int max_iters, int presteps,
                              double epslim, int errorswitch, tl_solve_info *info) {
  if (TL_IS_MULTI(c))
    return multi_solve(c, info, [&](tl_ctx *t, tl_solve_info *o, int) {
      return tl_cheby_solve(t, coef, rx, ry, eps, max_iters, presteps, epslim, errorswitch, o);
    });
  if (!c || !info) return TL_ERR_ARG;
  memset(info, 0, sizeof *info);
  CU(c, cudaSetDevice(c->device));
  max_iters = std::min(max_iters, c->max_iters);
  const long long l0 = c->launches;
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  const StopCfg cfg = switch_cfg(max_iters, eps, presteps, epslim, errorswitch);
  TRY(solve_preamble(c, coef, rx, ry, cfg));
  SolveState fin;
  TRY(cg_phase(c, &fin));
  TRY(cg_flush(c, fin.iter, true));
  CU(c, cudaEventRecord(c->ev_phase, c->stream));   // end of the CG phase (tl_get_option "last_cg_phase_ms")
  c->phase_marked = true;
  const int cgit = fin.iter;
  info->cg_iters = cgit;
  info->iters = cgit;
  info->error = cgit > 0 ? fin.red_rr : TL_ERROR_START;
  const bool converged = cgit > 0 && fabs(fin.red_rr) < eps;
  if (converged || cgit >= max_iters) { finish_timing(c, info, l0); return TL_OK; }

  // ---- switch: eigenvalues!, coef!  (Cheby.jl:66-67 with Appendix A #14) ----
  std::vector<double> al, be;
  TRY(fetch_cg_coefficients(c, cgit, al, be));
  double eigmin = 0, eigmax = 0;
  const int erc = tl::eigenvalues(al.data(), be.data(), cgit, &eigmin, &eigmax);
  info->eigmin = eigmin; info->eigmax = eigmax;
  if (erc) { finish_timing(c, info, l0); return tl_fail(c, TL_ERR_EIGEN, "Negative eigenvalue found: (%g, %g)", eigmin, eigmax); }
  const int ncoef = std::max(1, std::min(max_iters - cgit, c->max_iters));
  std::vector<double> cha(ncoef + 2, 0.0), chb(ncoef + 2, 0.0);
  const double theta = tl::cheby_coef(eigmin, eigmax, ncoef, cha.data(), chb.data());
  CU(c, cudaMemcpyAsync(c->ch_alphas, cha.data(), (ncoef + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->ch_betas, chb.data(), (ncoef + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  k_state_cheby<<<1, 1, 0, c->stream>>>(c->st, theta, eps, cgit + 1, max_iters);
  c->launches++;
  CHECK_LAUNCH(c);
  // both u buffers must agree outside the cells the fused kernel writes
  LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[TL_U], c->buf[B_U1]);
  // the first kernel pushes into the neighbours' B_U1 halos: their copy above must be done
  if (c->nranks > 1) TRY(tile_barrier(c));
  // Cheby.init! field part + bb
  TRY(launch_cheby<true>(c));
  c->launches++;
  if (legacy_comm(c)) TRY(allreduce2(c, &c->st->red_norm_local, &c->st->red_norm, 1));
  double bb = 0.0, error = 0.0;
  TRY(read_scalars(c, &c->st->red_norm, 1, &bb));
  // first main step with the norm, then Cheby.calciter
  TRY(enqueue_cheby_iteration(c));
  TRY(read_scalars(c, &c->st->red_norm, 1, &error));
  const int est = tl::cheby_calc_iter(eigmin, eigmax, error, bb);
  info->est_iters = est;
  k_state_set_est<<<1, 1, 0, c->stream>>>(c->st, est);
  c->launches++;
  CHECK_LAUNCH(c);
  auto enq = [&]() { return enqueue_cheby_iteration(c); };
  auto stop = [&](const SolveState &s) { return tl_cheby_should_stop(s); };
  bool done = false;
  const bool tiled_pairs = cheby_pairs_tiled(c);
  if ((c->cheby_pair && c->nranks == 1) || tiled_pairs) {
    // Norm iterations have odd tt (Cheby.jl:40-51: (tt+1) % 10 == 0), so pairs that start on an even tt
    // never have one as their first half: one single step aligns, then pairs run until the stop rule fires
    // or only one permitted iteration is left; single steps finish.
    const int tt_next = (cgit + 1) + 2 - 1;   // tt0 + cheby_step - 1 with cheby_step = 2 after the first main step
    if (tt_next & 1) TRY(enqueue_cheby_iteration(c));
    if (tiled_pairs) {
      const int step_now = 2 + (tt_next & 1);                       // cheby_step after the alignment step
      // kx, ky, u0, p and the buffer the first pair reads u from (no pair ran yet)
      const int bufs[5] = {TL_KX, TL_KY, TL_U0, TL_P, (step_now & 1) ? B_U1 : TL_U};
      TRY(pair_tiled_fill_halos(c, bufs, 5));
    }
    auto enq2 = [&]() { return enqueue_cheby_pair(c); };
    auto stop2 = [&](const SolveState &s) { return !cheby_pair_allowed(s); };
    TRY(run_chunks(c, &c->g_cheby2, &c->g_cheby2_iters, std::max(1, c->graph_iters / 2), 1, enq2, stop2, &fin));
    done = tl_cheby_should_stop(fin);
  }
  if (!done) TRY(run_chunks(c, &c->g_cheby, &c->g_cheby_iters, c->graph_iters, 1, enq, stop, &fin));
  c->u_cur = (fin.cheby_step + fin.cheby_pairs) & 1;
  c->p_cur = fin.cheby_pairs & 1;
  if (legacy_comm(c)) {   // haloupdate!(.., [:u]) of the last iteration on the tile-internal sides
    TRY(pull_halo(c, c->u_cur ? B_U1 : TL_U, 1));
    TRY(tile_barrier(c));
  }
  if (c->u_cur) {
    LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[B_U1], c->buf[TL_U]);
    c->u_cur = 0;
  }
  if (c->p_cur) {   // interior only: the Chebyshev iterations never touch p's halo (it stays as CG left it)
    LAUNCH_BASIC(c, k_copy, c->g, 0, c->buf[B_P1], c->buf[TL_P]);
    c->p_cur = 0;
  }
  finish_timing(c, info, l0);
  info->cheby_iters = fin.cheby_step - 1;
  info->iters = cgit + info->cheby_iters;
  info->error = fin.red_norm;
  return TL_OK;
}

static PpcgUrParams ppcg_ur_params(tl_ctx *c) {
  PpcgUrParams P;
  P.g = c->g; P.t = c->pw_tiling; P.st = c->st; P.hist_pw = c->hist_pw;
  P.p0 = c->buf[TL_P]; P.p1 = c->buf[B_P1]; P.w = c->buf[TL_W]; P.u = c->buf[TL_U]; P.r = c->buf[TL_R];
  P.sd0 = c->buf[TL_SD];
  P.partials = c->partials; P.cd = comm_dev(c); P.push_sd0 = push_for(c, TL_SD);
  P.deep = 0; P.r_out = P.r; P.d_sd = 1; P.d_r = 0;
  memset(&P.push_sd8, 0, sizeof P.push_sd8);
  memset(&P.push_r8, 0, sizeof P.push_r8);
  return P;
}
// the matrix-powers variant: r goes to the first group's input copy, halos k (sd) / k-1 (r) deep
static PpcgUrParams ppcg_ur_params_dk(tl_ctx *c, int inner_steps, int k) {
  PpcgUrParams P = ppcg_ur_params(c);
  const int G = (inner_steps + k - 1) / k, L0 = std::min(k, inner_steps);
  P.deep = 1;
  P.r_out = c->buf[(G & 1) ? B_R1 : TL_R];
  P.d_sd = L0; P.d_r = L0 - 1;
  P.push_sd8 = push8_for(c, TL_SD);
  P.push_r8 = push8_for(c, (G & 1) ? B_R1 : TL_R);
  return P;
}
static PpcgDkParams ppcg_dk_params(tl_ctx *c) {
  PpcgDkParams P;
  P.g = c->g; P.t = c->tiling; P.ext = c->dk_ext; P.st = c->st; P.alphas = c->ch_alphas; P.betas = c->ch_betas;
  P.sin0 = c->buf[TL_SD]; P.sin1 = c->buf[B_SD1]; P.wa = c->buf[B_SD2]; P.wb = c->buf[B_SD3];
  P.rin0 = c->buf[TL_R]; P.rin1 = c->buf[B_R1]; P.rw = c->buf[B_R2];
  P.u = c->buf[TL_U]; P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.partials = c->partials;
  P.cd = comm_dev(c); P.k = c->dk_k;
  P.push_sin0 = push8_for(c, TL_SD); P.push_sin1 = push8_for(c, B_SD1);
  P.push_rin0 = push8_for(c, TL_R); P.push_rin1 = push8_for(c, B_R1);
  return P;
}
template <int S, int MINB>
static int launch_ppcg_dk_ring(tl_ctx *c, const PpcgDkParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_ppcg_inner_dk<S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_ppcg_inner_dk<S, MINB>, c->dk_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static int launch_ppcg_dk(tl_ctx *c) {
  const PpcgDkParams P = ppcg_dk_params(c);
  switch (c->ring_eff) {
    case 3: TRY((launch_ppcg_dk_ring<3, 3>(c, P))); break;
    case 4: TRY((launch_ppcg_dk_ring<4, 2>(c, P))); break;
    case 6: TRY((launch_ppcg_dk_ring<6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  CHECK_LAUNCH(c);
  return TL_OK;
}
static PpcgInnerParams ppcg_inner_params(tl_ctx *c) {
  PpcgInnerParams P;
  P.g = c->g; P.t = c->tiling; P.st = c->st; P.alphas = c->ch_alphas; P.betas = c->ch_betas;
  P.sda = c->buf[TL_SD]; P.sdb = c->buf[B_SD1]; P.r = c->buf[TL_R]; P.u = c->buf[TL_U];
  P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.partials = c->partials;
  P.single = c->nranks == 1;
  P.cd = comm_dev(c); P.push_sda = push_for(c, TL_SD); P.push_sdb = push_for(c, B_SD1); P.push_r = push_for(c, TL_R);
  return P;
}

// One PPCG outer iteration.  Tiled (fused mode): every kernel pushes the operand the next kernel
// reads through the stencil (p'; sd0; sd'; r after the last inner step) and ends with the tile
// exchange.  Legacy mode: depth-1 halo pulls -- r, p before the matvec (ordered by the preceding rr
// allreduce) and sd before every inner step (ordered by a 1-double NCCL rendezvous).
// does this context run the inner steps two per pass?  Single tile: always (option ppcg_pair); tiles: when the
// pair kernels can run on them and the caller did not ask for a specific exchange depth (requested_depth = 0).
static bool ppcg_pairs_enabled(const tl_ctx *c, int inner_steps, int requested_depth) {
  if (!c->ppcg_pair || inner_steps < 2) return false;
  if (c->nranks == 1) return true;
  const int req = requested_depth > 0 ? requested_depth : c->ppcg_depth_k;
  return req == 0 && pairs_on_tiles(c);
}
template <int S, int MINB, bool TILED>
static int launch_ppcg_pair_ring(tl_ctx *c, const PpcgPairParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_ppcg_pair_ring<S, MINB, TILED>, smem, &prepared));
  CU(c, tl_launch(c, k_ppcg_pair_ring<S, MINB, TILED>, c->pair_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
// pair k of the npairs pairs of an outer iteration: sd alternates TL_SD / B_SD1 from TL_SD (where
// k_ppcg_ur_sd writes it); r alternates TL_R / B_R1 phased so that the LAST pair writes TL_R
static int ppcg_pair_rin(int k, int npairs) { return ((npairs - k) & 1) ? B_R1 : TL_R; }
static PpcgPairParams ppcg_pair_params(tl_ctx *c, int k, int npairs) {
  PpcgPairParams P;
  const int sout = (k & 1) ? TL_SD : B_SD1, rout = ppcg_pair_rin(k + 1, npairs);
  P.g = c->g; P.t = c->pair_tiling; P.st = c->st; P.alphas = c->ch_alphas; P.betas = c->ch_betas;
  P.sin = c->buf[(k & 1) ? B_SD1 : TL_SD];
  P.sout = c->buf[sout];
  P.rin = c->buf[ppcg_pair_rin(k, npairs)];
  P.rout = c->buf[rout];
  P.u = c->buf[TL_U]; P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.partials = c->partials;
  P.cd = comm_dev(c);
  P.push_s = push8_for(c, sout); P.push_r = push8_for(c, rout);
  return P;
}
static int launch_ppcg_pair(tl_ctx *c, int k, int npairs) {
  const PpcgPairParams P = ppcg_pair_params(c, k, npairs);
  if (c->pair_stages == 5) {
    if (c->nranks > 1) return launch_ppcg_pair_ring<5, 2, true>(c, P);
    return launch_ppcg_pair_ring<5, 2, false>(c, P);
  }
  if (c->nranks > 1) return launch_ppcg_pair_ring<4, 2, true>(c, P);
  return launch_ppcg_pair_ring<4, 2, false>(c, P);
}
// the single step that ends an odd number of inner steps: sd sits in the buffer npairs pairs left it in, r in TL_R
static PpcgInnerParams ppcg_trailing_params(tl_ctx *c, int npairs) {
  PpcgInnerParams P = ppcg_inner_params(c);
  const int cur = (npairs & 1) ? B_SD1 : TL_SD, oth = (npairs & 1) ? TL_SD : B_SD1;
  P.sda = c->buf[cur]; P.sdb = c->buf[oth];            // pp = inner_steps - 1 is even: reads sda, writes sdb
  P.push_sda = push_for(c, cur); P.push_sdb = push_for(c, oth);
  return P;
}
static int launch_ppcg_trailing(tl_ctx *c, int npairs) {
  const PpcgInnerParams P = ppcg_trailing_params(c, npairs);
  switch (c->ring_eff) {
    case 3: TRY((launch_ppcg_inner_ring<3, 3>(c, P))); break;
    case 4: TRY((launch_ppcg_inner_ring<4, 2>(c, P))); break;
    case 6: TRY((launch_ppcg_inner_ring<6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  return TL_OK;
}

static int prepare_ur_sd(tl_ctx *c) {
  static std::atomic<unsigned long long> prepared{0};
  return tl_prepare_smem(c, k_ppcg_ur_sd, 0, &prepared);
}
static int enqueue_ppcg_outer(tl_ctx *c, int inner_steps, int depth_k, bool pairs) {
  const bool legacy = legacy_comm(c);
  if (pairs) {
    const int npairs = inner_steps / 2;
    TRY(launch_cg_a<TL_U_NONE>(c));
    PpcgUrParams U = ppcg_ur_params(c);
    const int rin0 = ppcg_pair_rin(0, npairs);       // r goes where the first pair reads it
    if (c->nranks > 1) {                             // tiles: sd two cells / r one cell deep into the eight surrounding tiles
      U.deep = 1; U.r_out = c->buf[rin0]; U.d_sd = 2; U.d_r = 1;
      U.push_sd8 = push8_for(c, TL_SD); U.push_r8 = push8_for(c, rin0);
    } else if (rin0 != TL_R) { U.deep = 1; U.r_out = c->buf[rin0]; U.d_sd = 1; U.d_r = 0; }
    TRY(prepare_ur_sd(c));
    CU(c, tl_launch(c, k_ppcg_ur_sd, c->pw_grid, TL_FUSED_THREADS, 0, U));
    for (int k = 0; k < npairs; k++) TRY(launch_ppcg_pair(c, k, npairs));
    if (inner_steps & 1) TRY(launch_ppcg_trailing(c, npairs));
    CHECK_LAUNCH(c);
    c->launches += 2 + npairs + (inner_steps & 1);
    return TL_OK;
  }
  if (depth_k > 1) {
    // matrix-powers groups: one tile exchange per depth_k inner steps (PpcgDkParams)
    TRY(launch_cg_a<TL_U_NONE>(c));
    TRY(prepare_ur_sd(c));
    CU(c, tl_launch(c, k_ppcg_ur_sd, c->pw_grid, TL_FUSED_THREADS, 0, ppcg_ur_params_dk(c, inner_steps, depth_k)));
    for (int pp = 0; pp < inner_steps; pp++) TRY(launch_ppcg_dk(c));
    c->launches += 2 + inner_steps;
    return TL_OK;
  }
  if (legacy) {
    TRY(pull_halo(c, TL_R, 1));
    TRY(pull_halo(c, TL_P, 1));
    TRY(pull_halo(c, B_P1, 1));
  }
  TRY(launch_cg_a<TL_U_NONE>(c));
  if (legacy) TRY(allreduce2(c, &c->st->red_pw_local, &c->st->red_pw, 1));
  TRY(prepare_ur_sd(c));
  CU(c, tl_launch(c, k_ppcg_ur_sd, c->pw_grid, TL_FUSED_THREADS, 0, ppcg_ur_params(c)));
  for (int pp = 0; pp < inner_steps; pp++) {
    if (legacy) {
      TRY(tile_barrier(c));
      TRY(pull_halo(c, TL_SD, 1));
      TRY(pull_halo(c, B_SD1, 1));
    }
    TRY(launch_ppcg_inner(c));
  }
  if (legacy) TRY(allreduce2(c, &c->st->red_rr_local, &c->st->red_rr, 1));
  c->launches += 2 + inner_steps;
  return TL_OK;
}

// Exchange depth the inner steps of this context will use: the request (0 = the option
// "ppcg_halo_depth", whose 0 = halo_depth) limited to what the halo ring and the smallest tile
// can carry; 1 on a single tile and in the legacy communication mode.
static int ppcg_effective_depth(tl_ctx *c, int requested, int inner_steps) {
  if (c->nranks == 1 || !c->comm_fused) return 1;
  int k = requested > 0 ? requested : (c->ppcg_depth_k > 0 ? c->ppcg_depth_k : c->g.hd);
  k = std::min(k, c->g.hd);
  k = std::min(k, inner_steps);
  for (int r = 0; r < c->nranks; r++) k = std::min(k, std::min(c->rank_blob[r].nx, c->rank_blob[r].ny));
  return std::max(k, 1);
}

extern "C" int tl_ppcg_solve(tl_ctx *c, int coef, double rx, double ry, double eps, int max_iters, int presteps,
                             double epslim, int errorswitch, int inner_steps, int halo_depth_k, tl_solve_info *info) {
  if (TL_IS_MULTI(c))
    return multi_solve(c, info, [&](tl_ctx *t, tl_solve_info *o, int) {
      return tl_ppcg_solve(t, coef, rx, ry, eps, max_iters, presteps, epslim, errorswitch, inner_steps, halo_depth_k, o);
    });
  if (!c || !info || inner_steps < 1 || halo_depth_k < 0) return tl_fail(c, TL_ERR_ARG, "tl_ppcg_solve: bad argument");
  if (halo_depth_k > c->g.hd) return tl_fail(c, TL_ERR_ARG, "tl_ppcg_solve: halo_depth_k %d exceeds halo_depth %d", halo_depth_k, c->g.hd);
  if ((size_t)inner_steps + 2 > c->hist_len)   // the coefficient arrays were sized from max_iters at tl_create
    return tl_fail(c, TL_ERR_ARG, "tl_ppcg_solve: inner_steps %d does not fit the coefficient arrays (%zu entries, sized from max_iters)",
                   inner_steps, c->hist_len - 2);
  memset(info, 0, sizeof *info);
  CU(c, cudaSetDevice(c->device));
  if (c->nranks > 1 && !c->comm_ready) return tl_fail(c, TL_ERR_STATE, "tile context used before tl_comm_connect");
  // Inner-step schedule: two steps per pass (pairs; on tiles that is an exchange every 2 steps) unless switched off or
  // a specific exchange depth was requested; otherwise one kernel per step, grouped by the exchange depth on tiles.
  const bool pairs = ppcg_pairs_enabled(c, inner_steps, halo_depth_k);
  const int depth_k = pairs ? (c->nranks > 1 ? 2 : 1) : ppcg_effective_depth(c, halo_depth_k, inner_steps);
  const bool dk_mode = !pairs && depth_k > 1;
  info->halo_depth_k = depth_k;
  max_iters = std::min(max_iters, c->max_iters);
  const long long l0 = c->launches;
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  StopCfg cfg = switch_cfg(max_iters, eps, presteps, epslim, errorswitch);
  TRY(solve_preamble(c, coef, rx, ry, cfg));
  SolveState fin;
  TRY(cg_phase(c, &fin));
  TRY(cg_flush(c, fin.iter, true));
  CU(c, cudaEventRecord(c->ev_phase, c->stream));   // end of the CG phase (tl_get_option "last_cg_phase_ms")
  c->phase_marked = true;
  const int cgit = fin.iter;
  info->cg_iters = cgit;
  info->iters = cgit;
  info->error = cgit > 0 ? fin.red_rr : TL_ERROR_START;
  const bool converged = cgit > 0 && fabs(fin.red_rr) < eps;
  if (converged || cgit >= max_iters) { finish_timing(c, info, l0); return TL_OK; }

  // ---- switch (PPCG.jl:39-45 with Appendix A #14, #19) ----
  std::vector<double> al, be;
  TRY(fetch_cg_coefficients(c, cgit, al, be));
  double eigmin = 0, eigmax = 0;
  const int erc = tl::eigenvalues(al.data(), be.data(), cgit, &eigmin, &eigmax);
  info->eigmin = eigmin; info->eigmax = eigmax;
  if (erc) { finish_timing(c, info, l0); return tl_fail(c, TL_ERR_EIGEN, "Negative eigenvalue found: (%g, %g)", eigmin, eigmax); }
  std::vector<double> cha(inner_steps + 2, 0.0), chb(inner_steps + 2, 0.0);
  const double theta = tl::cheby_coef(eigmin, eigmax, inner_steps, cha.data(), chb.data());
  CU(c, cudaMemcpyAsync(c->ch_alphas, cha.data(), (inner_steps + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->ch_betas, chb.data(), (inner_steps + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TRY(residual_async(c));                                   // PPCG.jl:59
  TRY(halo_update_buf(c, TL_P, 1));                         // PPCG.jl:60
  TRY(norm2_async(c, TL_R, &c->st->red_rr));                // Appendix A #19: rro = sum r^2
  cfg.first_it = cgit;
  cfg.switch_min = INT_MAX;
  k_state_begin<<<1, 1, 0, c->stream>>>(c->st, cfg, cgit, theta, inner_steps);
  c->launches++;
  CHECK_LAUNCH(c);
  if (pairs && c->nranks > 1) {   // the pair kernels recompute cells in the tile-internal halos: kx, ky two deep there
    const int kbufs[2] = {TL_KX, TL_KY};
    TRY(pair_tiled_fill_halos(c, kbufs, 2));
  }
  if (dk_mode) {
    // the groups compute in the tile-internal halos: kx, ky are needed there, depth_k cells deep,
    // corner blocks included (CG.init! leaves them only partly defined, CG.jl:61-68)
    const int kbufs[2] = {TL_KX, TL_KY};
    TRY(tile_barrier(c));
    TRY(pull_halo_wide(c, kbufs, 2, depth_k));
    if (c->dk_k != depth_k || c->dk_grid == 0 || c->dk_rows_per_chunk != c->tiling.rows_per_chunk) {
      compute_dk_tiling(c, depth_k);
      if (c->g_ppcg) { cudaGraphExecDestroy(c->g_ppcg); c->g_ppcg = nullptr; }
    }
  }
  auto enq = [&]() { return enqueue_ppcg_outer(c, inner_steps, dk_mode ? depth_k : 1, pairs); };
  auto stop = [&](const SolveState &s) { return tl_should_stop(s.iter, s.red_rr, s.cfg); };
  const int chunk = std::max(1, c->graph_iters / 4);
  if (c->g_ppcg && (c->g_ppcg_inner != inner_steps || c->g_ppcg_k != (dk_mode ? depth_k : 1))) { cudaGraphExecDestroy(c->g_ppcg); c->g_ppcg = nullptr; }
  c->g_ppcg_inner = inner_steps;
  c->g_ppcg_k = dk_mode ? depth_k : 1;
  if (c->g_ppcg && c->g_ppcg_pairs != (int)pairs) { cudaGraphExecDestroy(c->g_ppcg); c->g_ppcg = nullptr; }
  c->g_ppcg_pairs = pairs;
  TRY(run_chunks(c, &c->g_ppcg, &c->g_ppcg_iters, chunk, pairs ? 2 + inner_steps / 2 + (inner_steps & 1) : 2 + inner_steps, enq, stop, &fin));
  TRY(cg_flush(c, fin.iter, false));
  // which buffer holds sd now: every pair / trailing step / group / step flips it
  c->sd_cur = (pairs ? inner_steps / 2 + (inner_steps & 1) : dk_mode ? (inner_steps + depth_k - 1) / depth_k : inner_steps) & 1;
  if (c->sd_cur) {
    LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[B_SD1], c->buf[TL_SD]);
    c->sd_cur = 0;
  }
  finish_timing(c, info, l0);
  info->cheby_iters = fin.iter - cgit;
  info->inner_total = info->cheby_iters * inner_steps;
  info->iters = fin.iter;
  info->error = fin.red_rr;
  return TL_OK;
}

// ---------------------------------------------------------------------------------------
// Jacobi (src/solvers/Jacobi.jl)
// ---------------------------------------------------------------------------------------
static int jacobi_init_async(tl_ctx *c, int coef, double rx, double ry) {
  if (coef < TL_CONDUCTIVITY) return tl_fail(c, TL_ERR_ARG, "Coefficient %d is not valid.", coef);  // Jacobi.jl:34-36
  c->p_cur = c->u_cur = c->sd_cur = 0;
  LAUNCH_BASIC(c, k_jacobi_init_fields, c->g, c->buf[TL_ENERGY], c->buf[TL_DENSITY], c->buf[TL_U], c->buf[TL_U0]);
  LAUNCH_BASIC(c, k_jacobi_init_k, c->g, coef, rx, ry, c->buf[TL_DENSITY], c->buf[TL_KX], c->buf[TL_KY]);
  return TL_OK;   // copyu! (Jacobi.jl:53) is the identity after u0 .= u
}

extern "C" int tl_jacobi_init(tl_ctx *c, int coef, double rx, double ry) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_jacobi_init(t, coef, rx, ry); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  TRY(jacobi_init_async(c, coef, rx, ry));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

extern "C" int tl_jacobi_iterate(tl_ctx *c, double *error) {
  if (TL_IS_MULTI(c)) return multi_scalar(c, error, [&](tl_ctx *t, double *o) { return tl_jacobi_iterate(t, o); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  LAUNCH_BASIC(c, k_copy, c->g, 1, field_ptr(c, TL_U), c->buf[TL_R]);          // Jacobi.jl:64  r .= u
  LAUNCH_BASIC(c, k_jacobi_sweep, c->g, c->buf[TL_U0], c->buf[TL_KX], c->buf[TL_KY], c->buf[TL_R], field_ptr(c, TL_U),
               c->partials, &c->st->counter, &c->st->red_aux[0]);
  TRY(allreduce(c, &c->st->red_aux[0], 1));
  double v;
  TRY(read_scalars(c, &c->st->red_aux[0], 1, &v));
  if (error) *error = v;
  return TL_OK;
}

static JacobiParams jacobi_params(tl_ctx *c, int force_resid) {
  JacobiParams P;
  P.g = c->g; P.t = c->tiling; P.st = c->st;
  P.u0 = c->buf[TL_U0]; P.ua = c->buf[TL_U]; P.ub = c->buf[B_U1]; P.r = c->buf[TL_R];
  P.kx = c->buf[TL_KX]; P.ky = c->buf[TL_KY]; P.partials = c->partials;
  P.single = c->nranks == 1; P.force_resid = force_resid;
  P.cd = comm_dev(c); P.push_ua = push_for(c, TL_U); P.push_ub = push_for(c, B_U1);
  return P;
}
template <int S, int MINB>
static int launch_jacobi_ring(tl_ctx *c, const JacobiParams &P) {
  const int smem = (TL_FUSED_THREADS / 32) * S * TL_RING_STAGE_BYTES;
  static std::atomic<unsigned long long> prepared{0};
  TRY(tl_prepare_smem(c, k_jacobi_fused_ring<S, MINB>, smem, &prepared));
  CU(c, tl_launch(c, k_jacobi_fused_ring<S, MINB>, c->fused_grid, TL_FUSED_THREADS, smem, P));
  return TL_OK;
}
static int launch_jacobi(tl_ctx *c) {
  const JacobiParams P = jacobi_params(c, 0);
  switch (c->ring_eff) {
    case 3: TRY((launch_jacobi_ring<3, 3>(c, P))); break;
    case 4: TRY((launch_jacobi_ring<4, 2>(c, P))); break;
    case 6: TRY((launch_jacobi_ring<6, 1>(c, P))); break;
    default: return tl_fail(c, TL_ERR_STATE, "internal: ring depth %d", c->ring_eff);
  }
  CHECK_LAUNCH(c);
  c->launches++;
  return TL_OK;
}
static int launch_jacobi_resid(tl_ctx *c, int force) {
  CU(c, tl_launch(c, k_jacobi_resid, c->basic_grid, TL_BASIC_THREADS, 0, jacobi_params(c, force)));
  c->launches++;
  return TL_OK;
}

// Jacobi.driver! (the module's solve!), src/solvers/Jacobi.jl:7-31.
// A graph holds 50 iteration kernels followed by the residual kernel of Jacobi.jl:16-21.
extern "C" int tl_jacobi_solve(tl_ctx *c, int coef, double rx, double ry, double eps, int max_iters, tl_solve_info *info) {
  if (TL_IS_MULTI(c))
    return multi_solve(c, info, [&](tl_ctx *t, tl_solve_info *o, int) { return tl_jacobi_solve(t, coef, rx, ry, eps, max_iters, o); });
  if (!c || !info) return TL_ERR_ARG;
  memset(info, 0, sizeof *info);
  CU(c, cudaSetDevice(c->device));
  if (legacy_comm(c)) return tl_fail(c, TL_ERR_STATE, "tl_jacobi_solve on tiles needs comm_fused = 1");
  max_iters = std::min(max_iters, c->max_iters);
  const long long l0 = c->launches;
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  TRY(jacobi_init_async(c, coef, rx, ry));
  LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[TL_U], c->buf[B_U1]);    // both u buffers agree in the deep halos
  StopCfg cfg{max_iters, TL_CONV_ABS, INT_MAX, 0, eps, 0.0};
  k_state_begin<<<1, 1, 0, c->stream>>>(c->st, cfg, 0, 0.0, 0);
  c->launches++;
  CHECK_LAUNCH(c);
  if (c->nranks > 1) TRY(tile_barrier(c));   // the first kernel pushes into the neighbours' B_U1 halos
  const long long saved = c->launches;
  auto enq = [&]() -> int {
    for (int i = 0; i < 50; i++) TRY(launch_jacobi(c));
    return launch_jacobi_resid(c, 0);
  };
  auto stop = [&](const SolveState &s) { return tl_should_stop(s.iter, s.red_rr, s.cfg); };
  SolveState fin;
  TRY(run_chunks(c, &c->g_jacobi, &c->g_jacobi_iters, 1, 51, enq, stop, &fin));
  (void)saved;
  // leave the reference's post-solve state: u in its own buffer, r = the previous iterate
  // (Jacobi.jl:64) or, after a 50th iteration, the residual on the interior
  const int cur = fin.iter & 1;
  LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[cur ? TL_U : B_U1], c->buf[TL_R]);
  if (fin.iter > 0 && fin.iter % 50 == 0) TRY(launch_jacobi_resid(c, 1));
  if (cur) LAUNCH_BASIC(c, k_copy, c->g, 1, c->buf[B_U1], c->buf[TL_U]);
  c->u_cur = 0;
  finish_timing(c, info, l0);
  info->iters = fin.iter;
  info->error = fin.iter > 0 ? fin.red_rr : TL_ERROR_START;
  return TL_OK;
}

// ---------------------------------------------------------------------------------------
// measurement helper
// ---------------------------------------------------------------------------------------
__global__ void k_state_for_timing(SolveState *st, double *hist_rr, double *hist_pw, double *cha, double *chb, int n, int iter) {
  StopCfg cfg{INT_MAX, TL_CONV_ABS, INT_MAX, 0, 0.0, 0.0};
  st->cfg = cfg;
  st->iter = iter;
  st->red_rr = 1.0; st->red_pw = 1e300;
  for (int i = 0; i < 5; i++) { hist_rr[i] = 1.0; hist_pw[i] = 1e300; }
  st->theta = 1.0; st->eps_cheby = 0.0; st->cheby_step = 1; st->cheby_pairs = 0; st->cheby_est = INT_MAX;
  st->cheby_tt0 = 1; st->cheby_max_tt = INT_MAX; st->inner_steps = INT_MAX; st->inner_pp = 0; st->counter = 0u;
  for (int i = 0; i < n; i++) { cha[i] = 0.5; chb[i] = 1e-3; }
}

extern "C" int tl_time_kernel(tl_ctx *c, const char *kernel, int reps, double *avg_ms) {
  if (TL_IS_MULTI(c)) return multi_max(c, avg_ms, [&](tl_ctx *t, double *o) { return tl_time_kernel(t, kernel, reps, o); });
  if (!c || !kernel || reps < 1 || !avg_ms) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  const std::string k(kernel);
  reps = std::min(reps, c->max_iters - 4);
  const int t_iter = (k == "cg_fused_w_odd") ? 3 : 2;
  k_state_for_timing<<<1, 1, 0, c->stream>>>(c->st, c->hist_rr, c->hist_pw, c->ch_alphas, c->ch_betas, c->max_iters, t_iter);
  CHECK_LAUNCH(c);
  auto launch = [&]() -> int {
    if (k == "cg_fused_w" || k == "cg_fused_w_odd") {
      // the kernel of the CG loop; lazy-u mode: iteration 2 applies both pending u updates, "_odd" (iteration 3) none
      if (cg_u_mode(c) != TL_U_LAZY) TRY(launch_cg_a<TL_U_EVERY>(c));
      else if (t_iter == 2) TRY(launch_cg_a<TL_U_LAZY>(c, 2));
      else TRY(launch_cg_a<TL_U_NONE>(c, 1));
    }
    else if (k == "cg_fused_w_nou") TRY(launch_cg_a<TL_U_NONE>(c));
    else if (k == "cg_fused_r") TRY(launch_cg_b(c));
    else if (k == "cheby_fused") TRY(launch_cheby<false>(c));
    else if (k == "cheby_pair") { TRY(enqueue_cheby_pair(c)); c->launches--; }
    else if (k == "ppcg_pair") TRY(launch_ppcg_pair(c, 0, 2));
    else if (k == "ppcg_inner") TRY(launch_ppcg_inner(c));
    else if (k == "jacobi_fused") { TRY(launch_jacobi(c)); c->launches--; }
    else return tl_fail(c, TL_ERR_ARG, "tl_time_kernel: unknown kernel %s", kernel);
    c->launches++;
    return TL_OK;
  };
  if (k == "cheby_pair") k_state_set_step<<<1, 1, 0, c->stream>>>(c->st, 2);   // step 1 is a norm iteration: pairs start at 2
  for (int i = 0; i < 3; i++) TRY(launch());
  CHECK_LAUNCH(c);
  if (k == "ppcg_pair") reps = std::min(reps, (c->max_iters - 16) / 2);
  if (k == "cheby_pair") { reps = std::min(reps, (c->max_iters - 16) / 2); k_state_set_step<<<1, 1, 0, c->stream>>>(c->st, 2); }
  if (k == "cg_fused_r") {  // B advances the iteration counter: rewind it
    k_state_for_timing<<<1, 1, 0, c->stream>>>(c->st, c->hist_rr, c->hist_pw, c->ch_alphas, c->ch_betas, c->max_iters, t_iter);
  }
  CU(c, cudaEventRecord(c->ev_start, c->stream));
  for (int i = 0; i < reps; i++) TRY(launch());
  CU(c, cudaEventRecord(c->ev_stop, c->stream));
  TRY(xchg_finalize(c));
  CU(c, cudaEventSynchronize(c->ev_stop));
  CHECK_LAUNCH(c);
  float ms = 0.f;
  CU(c, cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop));
  *avg_ms = ms / reps;
  return TL_OK;
}

extern "C" int tl_timer_start(tl_ctx *c) {
  if (TL_IS_MULTI(c)) return multi_all(c, [&](tl_ctx *t, int) { return tl_timer_start(t); });
  if (!c) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaEventRecord(c->ev_t0, c->stream));
  return TL_OK;
}

extern "C" int tl_timer_stop(tl_ctx *c, double *elapsed_ms) {
  if (TL_IS_MULTI(c)) return multi_max(c, elapsed_ms, [&](tl_ctx *t, double *o) { return tl_timer_stop(t, o); });
  if (!c || !elapsed_ms) return TL_ERR_ARG;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventRecord(c->ev_t1, c->stream));
  CU(c, cudaEventSynchronize(c->ev_t1));
  float ms = 0.f;
  CU(c, cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
  *elapsed_ms = ms;
  return TL_OK;
}

extern "C" int tl_launch_count(tl_ctx *c, long long *count) {
  if (TL_IS_MULTI(c) && count) return multi_launch_count(c, count);
  if (!c || !count) return TL_ERR_ARG;
  *count = c->launches;
  return TL_OK;
}

#include "tl_multi.inl"
