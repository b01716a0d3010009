// tl_multi.inl -- ONE context that spans several GPUs of ONE process (included at the end of tl_api.cu).
//
// The reference is one process with one `Chunk` (run.jl:45-47, src/TeaLeaf.jl:35-44): its host cannot drive one
// process per GPU without growing an MPI / Distributed layer it does not have.  tl_create_multi therefore returns a
// tl_ctx that looks like a single Chunk of the GLOBAL mesh to every entry point of the ABI and owns px x py tile
// contexts underneath (one per GPU, wired through the same peer-mapped slabs and mailboxes as the
// one-process-per-GPU mode; same-process tiles are addressed by plain pointers with peer access enabled).
//
// Why worker threads: the per-tile drivers block (they poll device state between graph launches and read scalars
// back), and the tiles' kernels wait for each other inside their tails, so the tiles' host sides must make progress
// concurrently.  Each tile gets one service thread inside the library; the caller stays single-threaded -- a call
// posts the same command to every tile and returns when all of them have finished (out-scalars are the all-tiles
// values every tile computed bit-identically; tile 0's copy is returned).
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct Multi {
  int n = 0, px = 1, py = 1;
  int gnx = 0, gny = 0, hd = 0;
  std::vector<tl_ctx *> tiles;
  std::vector<int> x0, y0, nx, ny;
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int)> job;
  unsigned long gen = 0;
  int pending = 0;
  bool quit = false;
  std::vector<int> rc;
};

static void multi_worker(Multi *m, int idx) {
  unsigned long seen = 0;
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_job.wait(lk, [&] { return m->quit || m->gen != seen; });
      if (m->quit) return;
      seen = m->gen;
      job = m->job;
    }
    const int r = job(idx);
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->rc[idx] = r;
      if (--m->pending == 0) m->cv_done.notify_all();
    }
  }
}

// run job(idx) on every tile's service thread; returns the first failing tile's status (its message becomes the
// multi context's last error)
static int multi_run(tl_ctx *c, const std::function<int(int)> &job) {
  Multi *m = c->multi;
  {
    std::unique_lock<std::mutex> lk(m->mu);
    m->job = job;
    m->pending = m->n;
    m->gen++;
    m->cv_job.notify_all();
    m->cv_done.wait(lk, [&] { return m->pending == 0; });
  }
  int first = TL_OK;
  std::string all;
  for (int i = 0; i < m->n; i++)
    if (m->rc[i] != TL_OK) {      // every failing tile's own story (a rendezvous failure is only understood from all of them)
      if (first == TL_OK) first = m->rc[i];
      all += (all.empty() ? "" : "; ") + ("tile " + std::to_string(i) + ": ") + (m->tiles[i] ? m->tiles[i]->err : std::string("not created"));
    }
  if (first != TL_OK) c->err = all;
  return first;
}

template <typename F>
static int multi_all(tl_ctx *c, F fn) {
  Multi *m = c->multi;
  return multi_run(c, [&](int i) { return fn(m->tiles[i], i); });
}

// fn(tile, &value): every tile computes the same all-tiles value; tile 0's is returned
template <typename F>
static int multi_scalar(tl_ctx *c, double *out, F fn) {
  Multi *m = c->multi;
  std::vector<double> tmp(m->n, out ? *out : 0.0);
  const int rc = multi_run(c, [&](int i) { return fn(m->tiles[i], &tmp[i]); });
  if (rc == TL_OK && out) *out = tmp[0];
  return rc;
}

// fn(tile, &value): the maximum over the tiles (device times)
template <typename F>
static int multi_max(tl_ctx *c, double *out, F fn) {
  Multi *m = c->multi;
  std::vector<double> tmp(m->n, 0.0);
  const int rc = multi_run(c, [&](int i) { return fn(m->tiles[i], &tmp[i]); });
  if (rc == TL_OK && out) *out = *std::max_element(tmp.begin(), tmp.end());
  return rc;
}

// fn(tile, &info, idx): a whole solve on every tile; iteration counts and the error are identical on all tiles
// (the stop rule is evaluated on all-reduced values), solve_ms is the slowest tile's, launches are summed
template <typename F>
static int multi_solve(tl_ctx *c, tl_solve_info *info, F fn) {
  Multi *m = c->multi;
  if (!info) return TL_ERR_ARG;
  std::vector<tl_solve_info> infos(m->n);
  const int rc = multi_run(c, [&](int i) { return fn(m->tiles[i], &infos[i], i); });
  *info = infos[0];
  for (int i = 1; i < m->n; i++) {
    info->solve_ms = std::max(info->solve_ms, infos[i].solve_ms);
    info->kernel_launches += infos[i].kernel_launches;
  }
  return rc;
}

// fn(tile, &value): the minimum over the tiles (health checks)
template <typename F>
static int multi_min(tl_ctx *c, double *out, F fn) {
  Multi *m = c->multi;
  std::vector<double> tmp(m->n, 0.0);
  const int rc = multi_run(c, [&](int i) { return fn(m->tiles[i], &tmp[i]); });
  if (rc == TL_OK && out) *out = *std::min_element(tmp.begin(), tmp.end());
  return rc;
}

static int multi_tile_offset(const tl_ctx *c, int idx, int *x0, int *y0) {
  *x0 = c->multi->x0[idx];
  *y0 = c->multi->y0[idx];
  return TL_OK;
}

static void multi_split(int n, int parts, int idx, int *off, int *size) {
  const int base = n / parts, rem = n % parts;
  *size = base + (idx < rem ? 1 : 0);
  *off = idx * base + std::min(idx, rem);
}

static void multi_destroy(tl_ctx *c) {
  Multi *m = c->multi;
  if (m->n && !m->workers.empty()) {
    // all tiles idle first (no kernel of one tile may address the freed memory of another), then free
    multi_run(c, [&](int i) {
      if (m->tiles[i]) { cudaSetDevice(m->tiles[i]->device); cudaStreamSynchronize(m->tiles[i]->stream); }
      return TL_OK;
    });
    multi_run(c, [&](int i) { if (m->tiles[i]) { tl_destroy(m->tiles[i]); m->tiles[i] = nullptr; } return TL_OK; });
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->quit = true;
      m->cv_job.notify_all();
    }
    for (auto &t : m->workers) t.join();
  }
  delete m;
  delete c;
}

extern "C" int tl_create_multi(tl_ctx **out, int xcells, int ycells, int halo_depth, int max_iters, int ngpus,
                               const int *devices, int px, int py) {
  if (!out) return TL_ERR_ARG;
  *out = nullptr;
  if (ngpus < 1 || ngpus > TL_MAX_RANKS || xcells < 1 || ycells < 1 || halo_depth < 1 || max_iters < 1) return TL_ERR_ARG;
  if (px <= 0 || py <= 0) {   // 1x1, 1x2, 2x2, 2x4 (y, the strided dimension, is split first), else the squarest split
    px = 1;
    for (int q = 1; q * q <= ngpus; q++)
      if (ngpus % q == 0) px = q;
    py = ngpus / px;
  }
  if (px * py != ngpus || px > xcells || py > ycells) {
    set_create_error("tl_create_multi: px*py must equal ngpus and every tile needs at least one cell");
    return TL_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return TL_ERR_NO_DEVICE; }
  std::vector<int> dev(ngpus);
  bool shared = false;
  for (int i = 0; i < ngpus; i++) {
    dev[i] = devices ? devices[i] : i;
    if (dev[i] < 0 || dev[i] >= ndev) { set_create_error("tl_create_multi: device index out of range"); return TL_ERR_NO_DEVICE; }
    for (int q = 0; q < i; q++) shared = shared || dev[q] == dev[i];
  }
  if (shared) {
    // several tiles on ONE GPU (a test mode: they share its bandwidth) wait for each other inside their kernels, so
    // nothing may synchronise the device while they run -- lazy module loading does, on a kernel's first launch
    const char *ml = getenv("CUDA_MODULE_LOADING");
    if (!ml || std::string(ml) != "EAGER") {
      set_create_error("tl_create_multi: tiles that share a GPU need CUDA_MODULE_LOADING=EAGER (and CUDA_DEVICE_MAX_CONNECTIONS >= "
                       "the number of tiles) set before CUDA initialises");
      return TL_ERR_STATE;
    }
  }
  tl_ctx *c = new tl_ctx();
  Multi *m = new Multi();
  c->multi = m;
  m->n = ngpus; m->px = px; m->py = py; m->gnx = xcells; m->gny = ycells; m->hd = halo_depth;
  c->g.nx = xcells; c->g.ny = ycells; c->g.hd = halo_depth; c->max_iters = max_iters;
  c->px = px; c->py = py; c->nranks = ngpus; c->device = dev[0];
  m->tiles.assign(ngpus, nullptr);
  m->rc.assign(ngpus, TL_OK);
  m->x0.resize(ngpus); m->y0.resize(ngpus); m->nx.resize(ngpus); m->ny.resize(ngpus);
  for (int r = 0; r < ngpus; r++) {
    multi_split(xcells, px, r % px, &m->x0[r], &m->nx[r]);
    multi_split(ycells, py, r / px, &m->y0[r], &m->ny[r]);
  }
  for (int r = 0; r < ngpus; r++) m->workers.emplace_back(multi_worker, m, r);
  int rc = multi_run(c, [&](int i) {
    return tl_create_tile(&m->tiles[i], m->nx[i], m->ny[i], halo_depth, max_iters, dev[i], i, px, py);
  });
  if (rc == TL_OK && ngpus > 1) {
    std::vector<CommBlob> blobs(ngpus);
    rc = multi_run(c, [&](int i) { return tl_comm_export(m->tiles[i], &blobs[i]); });
    if (rc == TL_OK) rc = multi_run(c, [&](int i) { return tl_comm_connect(m->tiles[i], blobs.data(), nullptr); });
  } else if (rc == TL_OK) {
    rc = tl_comm_connect(m->tiles[0], nullptr, nullptr);
  }
  if (rc != TL_OK) {
    set_create_error(c->err.empty() ? g_create_error : c->err);
    multi_destroy(c);
    return rc;
  }
  *out = c;
  return TL_OK;
}

// ---- scatter / gather of the reference's host matrices (global (x, y) array, halos included, leading dimension ld) ----
// A tile's padded array starts at global element (x0, y0): uploads copy the whole padded window (tile-internal halos
// receive the neighbours' cells, exactly what an exchange would put there); downloads copy each tile's interior plus
// the halo sides that are physical, so every host element is written by exactly one tile.
static int get_field_rect(tl_ctx *c, int field, double *host, long ld, int xl, int xr, int yl, int yr) {
  const Geo &g = c->g;
  CU(c, cudaSetDevice(c->device));
  const double *src = field_ptr(c, field) - (long)g.hd * g.pitch - g.hd + (long)yl * g.pitch + xl;
  CU(c, cudaMemcpy2DAsync(host + (long)yl * ld + xl, (size_t)ld * 8, src, (size_t)g.pitch * 8, (size_t)(xr - xl) * 8, yr - yl,
                          cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return TL_OK;
}

static int multi_set_field(tl_ctx *c, int field, const double *host, long ld) {
  Multi *m = c->multi;
  if (!host || field < 0 || field >= TL_NUM_FIELDS || ld < m->gnx + 2 * m->hd) return tl_fail(c, TL_ERR_ARG, "tl_set_field: bad argument");
  return multi_run(c, [&](int i) { return tl_set_field(m->tiles[i], field, host + (long)m->y0[i] * ld + m->x0[i], ld); });
}

static int multi_get_field(tl_ctx *c, int field, double *host, long ld) {
  Multi *m = c->multi;
  if (!host || field < 0 || field >= TL_NUM_FIELDS || ld < m->gnx + 2 * m->hd) return tl_fail(c, TL_ERR_ARG, "tl_get_field: bad argument");
  return multi_run(c, [&](int i) {
    tl_ctx *t = m->tiles[i];
    const int hd = m->hd, cx = i % m->px, cy = i / m->px;
    const int xl = cx == 0 ? 0 : hd, xr = m->nx[i] + (cx == m->px - 1 ? 2 * hd : hd);
    const int yl = cy == 0 ? 0 : hd, yr = m->ny[i] + (cy == m->py - 1 ? 2 * hd : hd);
    return get_field_rect(t, field, host + (long)m->y0[i] * ld + m->x0[i], ld, xl, xr, yl, yr);
  });
}

static int multi_field_summary(tl_ctx *c, double cell_volume, double *vol, double *mass, double *ie, double *temp) {
  Multi *m = c->multi;
  std::vector<double> v(4 * m->n, 0.0);
  const int rc = multi_run(c, [&](int i) {
    return tl_field_summary(m->tiles[i], cell_volume, &v[4 * i], &v[4 * i + 1], &v[4 * i + 2], &v[4 * i + 3]);
  });
  if (rc != TL_OK) return rc;
  if (vol) *vol = v[0];
  if (mass) *mass = v[1];
  if (ie) *ie = v[2];
  if (temp) *temp = v[3];
  return TL_OK;
}

static int multi_launch_count(tl_ctx *c, long long *count) {
  Multi *m = c->multi;
  long long total = 0;
  for (int i = 0; i < m->n; i++) total += m->tiles[i]->launches;
  *count = total;
  return TL_OK;
}
