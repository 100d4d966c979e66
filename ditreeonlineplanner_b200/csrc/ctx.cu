// ctx.cu -- context lifetime, map upload, status word.
#include "carfast.cuh"

void dt_denoiser_free(dt_ctx* ctx);         // denoiser.cu
void dt_denoiser_drop_graphs(dt_ctx* ctx);  // denoiser.cu

extern "C" const char* dt_version(void) { return "ditree-b200 0.1 (sm_100a)"; }

extern "C" int dt_ctx_create(int device, dt_ctx** out) {
  if (!out) return DT_E_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || device < 0 || device >= count) return DT_E_CUDA;
  dt_ctx* ctx = new dt_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DT_E_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DT_E_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  {
    const char* e = getenv("DITREE_PDL");
    if (e && e[0] == '0') ctx->pdl_on = false;
    e = getenv("DITREE_FORK");
    if (e && e[0] == '0') ctx->fork_on = false;
  }
  if (prop.major != 10) {
    // sm_100a cubins only load on compute capability 10.x parts; fail loudly instead of later
    fprintf(stderr, "libditree: device %d is sm_%d%d, this library is built for sm_100a only\n", device, prop.major,
            prop.minor);
    delete ctx;
    return DT_E_UNSUPPORTED;
  }
  if (cudaMalloc(&ctx->d_status, sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_status, 0, sizeof(int)) != cudaSuccess ||
      cudaMallocHost(&ctx->h_status, sizeof(int)) != cudaSuccess) {
    delete ctx;
    return DT_E_CUDA;
  }
  *out = ctx;
  return DT_OK;
}

extern "C" void dt_ctx_destroy(dt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  dt_denoiser_free(ctx);
  if (ctx->d_map) cudaFree(ctx->d_map);
  if (ctx->d_qmap) cudaFree(ctx->d_qmap);
  for (dt_map_slot& sl : ctx->slots) {
    if (sl.d_map) cudaFree(sl.d_map);
    if (sl.d_qmap) cudaFree(sl.d_qmap);
  }
  if (ctx->d_map_table) cudaFree(ctx->d_map_table);
  if (ctx->d_status) cudaFree(ctx->d_status);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_wide) cudaFree(ctx->d_wide);
  if (ctx->d_splitk) cudaFree(ctx->d_splitk);
  if (ctx->d_splitk2) cudaFree(ctx->d_splitk2);
  for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
  delete ctx;
}

extern "C" int dt_set_option(dt_ctx* ctx, const char* name, int value) {
  if (!ctx || !name) return DT_E_ARG;
  if (strcmp(name, "splitk") == 0) {
    if (ctx->splitk_on != (value != 0)) {
      DT_CUDA(cudaSetDevice(ctx->device));
      DT_CUDA(cudaDeviceSynchronize());  // no replay of a graph that is about to be destroyed may be in flight
      dt_denoiser_drop_graphs(ctx);
    }
    ctx->splitk_on = value != 0;
    return DT_OK;
  }
  if (strcmp(name, "fork") == 0) {
    if (ctx->fork_on != (value != 0)) {  // captured graphs hold the fork / join edges
      DT_CUDA(cudaSetDevice(ctx->device));
      DT_CUDA(cudaDeviceSynchronize());
      dt_denoiser_drop_graphs(ctx);
    }
    ctx->fork_on = value != 0;
    return DT_OK;
  }
  if (strcmp(name, "pdl") == 0) {
    if (ctx->pdl_on != (value != 0)) {  // captured graphs hold the launch attributes of their kernel nodes
      DT_CUDA(cudaSetDevice(ctx->device));
      DT_CUDA(cudaDeviceSynchronize());
      dt_denoiser_drop_graphs(ctx);
    }
    ctx->pdl_on = value != 0;
    return DT_OK;
  }
  return dt_fail(ctx, DT_E_ARG, "dt_set_option: unknown option");
}

extern "C" const char* dt_last_error(dt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int dt_profile_begin(dt_ctx* ctx) {
  if (!ctx) return DT_E_ARG;
  ctx->prof_on = true;
  ctx->prof_used = 0;
  ctx->prof_recs.clear();
  return DT_OK;
}

extern "C" int dt_profile_end(dt_ctx* ctx, double* gemm_ms_out, int64_t* gemm_launches_out) {
  if (!ctx) return DT_E_ARG;
  ctx->prof_on = false;
  DT_CUDA(cudaDeviceSynchronize());
  double ms = 0.0;
  for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
    float t = 0.f;
    DT_CUDA(cudaEventElapsedTime(&t, ctx->prof_events[i], ctx->prof_events[i + 1]));
    ms += t;
    if (i / 2 < ctx->prof_recs.size()) ctx->prof_recs[i / 2].ms = t;
  }
  if (gemm_ms_out) *gemm_ms_out = ms;
  if (gemm_launches_out) *gemm_launches_out = (int64_t)(ctx->prof_used / 2);
  ctx->prof_used = 0;
  return DT_OK;
}

extern "C" int dt_profile_csv(dt_ctx* ctx, const char* path) {
  if (!ctx || !path) return DT_E_ARG;
  FILE* f = fopen(path, "w");
  if (!f) return dt_fail(ctx, DT_E_ARG, "dt_profile_csv: cannot open file");
  fprintf(f, "idx,BN,epi,gw,M,N,K,ms,tflops,ksplit\n");
  for (size_t i = 0; i < ctx->prof_recs.size(); ++i) {
    const dt_ctx::ProfRec& r = ctx->prof_recs[i];
    const double fl = 2.0 * (double)r.M * r.N * (double)r.K;
    fprintf(f, "%zu,%d,%d,%d,%lld,%d,%lld,%.5f,%.1f,%d\n", i, r.bn, r.epi, r.gw, r.M, r.N, r.K, r.ms,
            (r.ms > 0 && r.epi != 2) ? fl / (r.ms * 1e-3) / 1e12 : 0.0, r.ksplit);
  }
  fclose(f);
  return DT_OK;
}

// next event of the profiling pool (created on demand)
cudaEvent_t dt_prof_event(dt_ctx* ctx) {
  if (ctx->prof_used == ctx->prof_events.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    ctx->prof_events.push_back(e);
  }
  return ctx->prof_events[ctx->prof_used++];
}

extern "C" int64_t dt_launch_count(dt_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dt_ensure_scratch(dt_ctx* ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return DT_OK;
  if (ctx->d_scratch) {
    DT_CUDA(cudaDeviceSynchronize());
    DT_CUDA(cudaFree(ctx->d_scratch));
    ctx->d_scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  size_t sz = bytes < (1u << 20) ? (1u << 20) : bytes;
  DT_CUDA(cudaMalloc(&ctx->d_scratch, sz));
  ctx->scratch_bytes = sz;
  return DT_OK;
}

// Host-side form of a grid: bytes (1 = wall, 2 = other non-zero, 0 = free; padded to 16) and the quadrant table of
// the car collision fast path.
static void dt_build_map_host(const float* grid_host, int rows, int cols, std::vector<uint8_t>& bytes,
                              std::vector<uint32_t>& q, int& padded, int& qpadded) {
  padded = ((rows * cols + 15) / 16) * 16;
  bytes.assign(padded, 0);
  for (int i = 0; i < rows * cols; ++i) bytes[i] = (grid_host[i] == 1.0f) ? 1 : (grid_host[i] != 0.0f ? 2 : 0);
  // Quadrant table of the car collision fast path (carfast.cuh).  The grid is padded by DT_QPAD rings of
  // "always collides" cells (a ball centre outside the grid collides, common/map_utils.py:255-259) and the
  // table is indexed by the grid VERTEX (k, j) nearest to the ball centre; each vertex holds four 32-bit
  // words, one per cell quadrant touching it: word 2a + b describes the cell above (a = 1) / below (a = 0)
  // and left (b = 1) / right (b = 0) of the vertex, i.e. the ball sits in the lower (a = 1) / upper half and
  // right (b = 1) / left half of that cell.  The flags sit at the sign-bit position of one byte each, so
  // the kernel can AND them with the gathered top bytes of its signed margins:
  //   bit  7  the side neighbour in x on that half is a wall   (index clipped as map_utils.py:281-296)
  //   bit 15  the side neighbour in y on that half is a wall
  //   bit 23  the diagonal neighbour of that quadrant is a wall (row clipped with R-1, column with R-1: sic, :326)
  //   bit 31  collides wherever the ball is: padding cell, own cell is a wall, or any of the four diagonal
  //           cells lies outside the grid (`invalid_cell`, :322-327, which makes every border cell collide)
  // Built only when rows <= cols (on taller maps the row-count clip can index past the last column: NumPy
  // raises IndexError, and the exact code, which reports that, is used instead) and when it fits in
  // DT_QMAP_MAX_BYTES of shared memory.
  q.clear();
  qpadded = 0;
  const int VR = rows + 2 * DT_QPAD + 1, VC = cols + 2 * DT_QPAD + 1;
  if (rows <= cols && (size_t)VR * VC * 16 <= DT_QMAP_MAX_BYTES) {
    qpadded = VR * VC * 16;
    q.assign((size_t)VR * VC * 4, 0x80000000u);
    auto wall = [&](int r, int c) { return bytes[r * cols + c] == 1; };
    auto clip = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
    for (int k = 0; k < VR; ++k)
      for (int j = 0; j < VC; ++j)
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            const int r = k - a - DT_QPAD, c = j - b - DT_QPAD;  // the cell of this quadrant, grid indices
            if (r < 0 || r >= rows || c < 0 || c >= cols) continue;
            const bool border = (r == 0) | (r == rows - 1) | (c == 0) | (c == cols - 1);  // a diagonal is outside
            uint32_t w = (wall(r, c) || border) ? 0x80000000u : 0u;
            const int sx = b ? 1 : -1, sy = a ? 1 : -1;
            w |= wall(r, clip(c + sx, 0, cols - 1)) ? 0x80u : 0u;
            w |= wall(clip(r + sy, 0, rows - 1), c) ? 0x8000u : 0u;
            w |= wall(clip(r + sy, 0, rows - 1), clip(c + sx, 0, rows - 1)) ? 0x800000u : 0u;  // sic: rows - 1
            q[((size_t)k * VC + j) * 4 + 2 * a + b] = w;
          }
  }
}

extern "C" int dt_set_map(dt_ctx* ctx, const float* grid_host, int rows, int cols, float s_global, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!grid_host || rows < 1 || cols < 1 || (int64_t)rows * cols > DT_MAX_MAP_CELLS || !(s_global > 0.f))
    return dt_fail(ctx, DT_E_ARG, "dt_set_map: bad grid (1 <= rows*cols <= 16384, s_global > 0)");
  DT_CUDA(cudaSetDevice(ctx->device));
  std::vector<uint8_t> bytes;
  std::vector<uint32_t> q;
  int padded = 0, qpadded = 0;
  dt_build_map_host(grid_host, rows, cols, bytes, q, padded, qpadded);
  cudaStream_t st = (cudaStream_t)stream;
  if ((size_t)padded > ctx->map_capacity || (size_t)qpadded > ctx->qmap_capacity) {
    DT_CUDA(cudaStreamSynchronize(st));
    if ((size_t)padded > ctx->map_capacity) {
      if (ctx->d_map) DT_CUDA(cudaFree(ctx->d_map));
      ctx->d_map = nullptr;
      ctx->map_capacity = 0;
      DT_CUDA(cudaMalloc(&ctx->d_map, padded));
      ctx->map_capacity = padded;
    }
    if ((size_t)qpadded > ctx->qmap_capacity) {
      if (ctx->d_qmap) DT_CUDA(cudaFree(ctx->d_qmap));
      ctx->d_qmap = nullptr;
      ctx->qmap_capacity = 0;
      DT_CUDA(cudaMalloc(&ctx->d_qmap, qpadded));
      ctx->qmap_capacity = qpadded;
    }
  }
  // synchronous: the staging vectors die at return, and the reference's update_maze is synchronous too
  DT_CUDA(cudaMemcpyAsync(ctx->d_map, bytes.data(), padded, cudaMemcpyHostToDevice, st));
  if (qpadded) DT_CUDA(cudaMemcpyAsync(ctx->d_qmap, q.data(), qpadded, cudaMemcpyHostToDevice, st));
  DT_CUDA(cudaStreamSynchronize(st));
  ctx->qmap_bytes = qpadded;
  ctx->rows = rows;
  ctx->cols = cols;
  ctx->s_global = (double)s_global;
  ctx->map_bytes = padded;
  return DT_OK;
}

extern "C" int dt_set_map_slot(dt_ctx* ctx, int slot, const float* grid_host, int rows, int cols, float s_global,
                               void* stream) {
  if (!ctx) return DT_E_ARG;
  if (slot < 0 || slot >= DT_MAX_MAP_SLOTS) return dt_fail(ctx, DT_E_ARG, "dt_set_map_slot: slot out of range (0..31)");
  if (!grid_host || rows < 1 || cols < 1 || (int64_t)rows * cols > DT_MAX_MAP_CELLS || !(s_global > 0.f))
    return dt_fail(ctx, DT_E_ARG, "dt_set_map_slot: bad grid (1 <= rows*cols <= 16384, s_global > 0)");
  DT_CUDA(cudaSetDevice(ctx->device));
  std::vector<uint8_t> bytes;
  std::vector<uint32_t> q;
  int padded = 0, qpadded = 0;
  dt_build_map_host(grid_host, rows, cols, bytes, q, padded, qpadded);
  cudaStream_t st = (cudaStream_t)stream;
  dt_map_slot& s = ctx->slots[slot];
  DT_CUDA(cudaStreamSynchronize(st));  // a pass still reading the old contents of this slot must have drained
  if ((size_t)padded > s.map_cap) {
    if (s.d_map) DT_CUDA(cudaFree(s.d_map));
    s.d_map = nullptr; s.map_cap = 0;
    DT_CUDA(cudaMalloc(&s.d_map, padded));
    s.map_cap = padded;
  }
  if ((size_t)qpadded > s.qmap_cap) {
    if (s.d_qmap) DT_CUDA(cudaFree(s.d_qmap));
    s.d_qmap = nullptr; s.qmap_cap = 0;
    DT_CUDA(cudaMalloc(&s.d_qmap, qpadded));
    s.qmap_cap = qpadded;
  }
  if (!ctx->d_map_table) {
    DT_CUDA(cudaMalloc(&ctx->d_map_table, sizeof(MapEntry) * DT_MAX_MAP_SLOTS));
    DT_CUDA(cudaMemset(ctx->d_map_table, 0, sizeof(MapEntry) * DT_MAX_MAP_SLOTS));
  }
  DT_CUDA(cudaMemcpyAsync(s.d_map, bytes.data(), padded, cudaMemcpyHostToDevice, st));
  if (qpadded) DT_CUDA(cudaMemcpyAsync(s.d_qmap, q.data(), qpadded, cudaMemcpyHostToDevice, st));
  s.rows = rows; s.cols = cols; s.map_bytes = padded; s.qmap_bytes = qpadded; s.s_global = (double)s_global;
  MapEntry e;
  e.m.g = s.d_map; e.m.rows = rows; e.m.cols = cols; e.m.bytes = padded; e.m.s = (double)s_global;
  e.q = dt_qmap_view_of(s.d_qmap, qpadded, rows, cols);
  DT_CUDA(cudaMemcpyAsync((MapEntry*)ctx->d_map_table + slot, &e, sizeof e, cudaMemcpyHostToDevice, st));
  DT_CUDA(cudaStreamSynchronize(st));  // the staging vectors and `e` die at return
  if (padded + qpadded > ctx->slots_max_bytes) ctx->slots_max_bytes = padded + qpadded;
  return DT_OK;
}

extern "C" int dt_sync_status(dt_ctx* ctx, void* stream) {
  if (!ctx) return DT_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  DT_CUDA(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaMemsetAsync(ctx->d_status, 0, sizeof(int), st));
  DT_CUDA(cudaStreamSynchronize(st));
  const int s = *ctx->h_status;
  if (s == DT_E_INDEX) ctx->err = "index out of range (the reference raises IndexError here)";
  return s;
}
