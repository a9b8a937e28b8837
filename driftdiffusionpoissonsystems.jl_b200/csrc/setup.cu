// One-time pattern construction on the device: replaces every COO -> CSC `sparse(I,J,V)` call of the
// reference (src:275,284,361,391,418,550,559,562,584,1019,1028,1031,1062,1087) by a FIXED pattern
// plus row -> incident-triangle lists, so the per-iteration assembly is a pure gather.
//
// Library primitives (CUB radix sort / scan / unique / reduce from the CUDA toolkit) are used only
// here, in setup; the per-iteration kernels are hand-written (kernels.cu).
#include <cub/cub.cuh>

#include <algorithm>
#include <numeric>

#include "ddps_internal.cuh"

namespace ddps {

namespace {

__global__ void k_make_inc_pairs(const int4* __restrict__ elem, int nme, int* __restrict__ keys, int* __restrict__ vals) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nme) return;
  int4 el = elem[e];
  int64_t o = 3LL * e;
  keys[o + 0] = el.x; vals[o + 0] = e * 4 + 0;
  keys[o + 1] = el.y; vals[o + 1] = e * 4 + 1;
  keys[o + 2] = el.z; vals[o + 2] = e * 4 + 2;
}

__global__ void k_make_adj_keys(const int4* __restrict__ elem, int nme, int n_own, unsigned long long* __restrict__ keys) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < nme) {
    int4 el = elem[t];
    unsigned long long a = (unsigned)el.x, b = (unsigned)el.y, c = (unsigned)el.z;
    int64_t o = 6LL * t;
    keys[o + 0] = (a << 32) | b; keys[o + 1] = (a << 32) | c;
    keys[o + 2] = (b << 32) | a; keys[o + 3] = (b << 32) | c;
    keys[o + 4] = (c << 32) | a; keys[o + 5] = (c << 32) | b;
  } else if (t < (int64_t)nme + n_own) {
    unsigned long long i = (unsigned long long)(t - nme);
    keys[6LL * nme + (t - nme)] = (i << 32) | i;   // diagonal (identity rows of eliminated nodes, src:279-281)
  }
}

template <typename K>
__global__ void k_lower_bound(const K* __restrict__ sorted, int64_t n, int nq, int shift, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  K target = ((K)i) << shift;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < target) lo = mid + 1; else hi = mid;
  }
  out[i] = (int)lo;
}

__global__ void k_low32(const unsigned long long* __restrict__ keys, int64_t n, int* __restrict__ col) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) col[i] = (int)(keys[i] & 0xffffffffull);
}

// slice width (in entries = 32*max row length) from a CSR-like pointer array
__global__ void k_slice_width(const int* __restrict__ ptr, int nrows, int nslices, long long* __restrict__ width32,
                              int* __restrict__ maxlen) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = row >> 5;
  if (slice >= nslices) return;
  int len = (row < nrows) ? ptr[row + 1] - ptr[row] : 0;
  int m = len;
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((row & 31) == 0) {
    width32[slice] = 32LL * m;
    atomicMax(maxlen, m);
  }
}

__global__ void k_narrow(const long long* __restrict__ in, int n, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

__global__ void k_fill_sell_col(const int* __restrict__ rowptr, const int* __restrict__ csrcol, int nrows, int nslices,
                                const int* __restrict__ slice_ptr, int* __restrict__ col) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = row >> 5, lane = row & 31;
  if (slice >= nslices) return;
  int base = slice_ptr[slice];
  int w = (slice_ptr[slice + 1] - base) >> 5;
  int beg = 0, len = 0;
  if (row < nrows) { beg = rowptr[row]; len = rowptr[row + 1] - beg; }
  int self = row < nrows ? row : 0;
  for (int k = 0; k < w; ++k) col[base + k * 32 + lane] = (k < len) ? csrcol[beg + k] : self;
}

__device__ __forceinline__ int find_slot(const int* __restrict__ cols, int len, int target) {
  int lo = 0, hi = len;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cols[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// keys (slice << 32 | column) of every stored entry: sorted + uniqued they give, per slice of 32
// consecutive rows (= the warp that assembles them), the sorted list of distinct nodes its rows touch
__global__ void k_make_block_keys(const int* __restrict__ rowptr, const int* __restrict__ csrcol, int nrows,
                                  unsigned long long* __restrict__ keys) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const unsigned long long hi = (unsigned long long)(row / kSlice) << 32;
  for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) keys[k] = hi | (unsigned)csrcol[k];
}

__global__ void k_block_maxlen(const int* __restrict__ bn_ptr, int nblk, int* __restrict__ out_max) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nblk) atomicMax(out_max, bn_ptr[b + 1] - bn_ptr[b]);
}

// Incidence streams, slice-interleaved like the matrix (entry k of row r at inc_slice_ptr[r/32] + 32k + r%32):
//   inc_tri  = triangle*4 + corner (or -1 for padding)
//   inc_pack = slot_next | slot_prev<<5 | l_next<<10 | l_prev<<20 | again_next<<30 | again_prev<<31 (0xffffffff for
//              padding), slot_* = positions of the triangle's two off-diagonal entries in the row, l_* = positions of the
//              two other corners in the SLICE's node list (what the assembling warp stages in shared memory), again_* =
//              an earlier triangle of this row already contributed to that slot (so the first contribution is a plain
//              store and only the second one -- an edge has at most two triangles -- reads the accumulator back)
//   arinc    = signed area of the triangle in the row's own frame (A = the row's corner, B = next, C = previous corner,
//              src:243-246 up to the cyclic relabelling): the Newton-system kernels need nothing else of the geometry
__global__ void k_fill_incidence(const int* __restrict__ inc_ptr, const int* __restrict__ inc_vals,
                                 const int4* __restrict__ elem, const double2* __restrict__ xy,
                                 const int* __restrict__ rowptr, const int* __restrict__ csrcol, int nrows, int nslices,
                                 const int* __restrict__ inc_slice_ptr, const int* __restrict__ sn_ptr,
                                 const int* __restrict__ sn_nodes, int* __restrict__ inc_tri,
                                 unsigned* __restrict__ inc_pack, double* __restrict__ arinc,
                                 unsigned short* __restrict__ rowinfo) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  int slice = row >> 5, lane = row & 31;
  if (slice >= nslices) return;
  int base = inc_slice_ptr[slice];
  int w = (inc_slice_ptr[slice + 1] - base) >> 5;
  int beg = 0, len = 0, rbeg = 0, rlen = 0;
  const int lb = sn_ptr[slice], ll = sn_ptr[slice + 1] - lb;
  if (row < nrows) {
    beg = inc_ptr[row]; len = inc_ptr[row + 1] - beg;
    rbeg = rowptr[row]; rlen = rowptr[row + 1] - rbeg;
    rowinfo[row] = (unsigned short)(find_slot(csrcol + rbeg, rlen, row) | (rlen << 5));   // diagonal slot | row length
  }
  unsigned seen = 0;
  for (int k = 0; k < w; ++k) {
    int tri = -1;
    unsigned pack = 0xffffffffu;
    double ar = 0.0;
    if (k < len) {
      tri = inc_vals[beg + k];
      int e = tri >> 2, a = tri & 3;
      int4 el = elem[e];
      int nb = (a == 0) ? el.y : (a == 1) ? el.z : el.x;   // next corner
      int nc = (a == 0) ? el.z : (a == 1) ? el.x : el.y;   // previous corner
      unsigned sb = find_slot(csrcol + rbeg, rlen, nb), sc = find_slot(csrcol + rbeg, rlen, nc);
      unsigned l1 = find_slot(sn_nodes + lb, ll, nb), l2 = find_slot(sn_nodes + lb, ll, nc);
      pack = sb | (sc << 5) | (l1 << 10) | (l2 << 20) | (((seen >> sb) & 1u) << 30) | (((seen >> sc) & 1u) << 31);
      seen |= (1u << sb) | (1u << sc);
      const double2 qa = xy[row], qb = xy[nb], qc = xy[nc];
      const double eax = qb.x - qc.x, eay = qb.y - qc.y;   // e_A = q_B - q_C
      const double ebx = qc.x - qa.x, eby = qc.y - qa.y;   // e_B = q_C - q_A
      ar = (eax * eby - eay * ebx) * 0.5;
    }
    inc_tri[base + k * 32 + lane] = tri;
    inc_pack[base + k * 32 + lane] = pack;
    arinc[base + k * 32 + lane] = ar;
  }
}

// Runs of consecutive node ids in every slice's node list: the assembling warp copies a run with ONE bulk copy (its
// records are contiguous in memory).  Pass 1 (runs == nullptr) counts, pass 2 writes {first node, position in the list |
// length << 16} at run_ptr[slice].
__global__ void k_slice_runs(int nslices, const int* __restrict__ sn_ptr, const int* __restrict__ sn_nodes,
                             const int* __restrict__ run_ptr, int* __restrict__ counts, int2* __restrict__ runs) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int b = sn_ptr[s], e = sn_ptr[s + 1];
  int n = 0, start = b;
  int2* out = runs ? runs + run_ptr[s] : nullptr;
  for (int i = b + 1; i <= e; ++i) {
    if (i == e || sn_nodes[i] != sn_nodes[i - 1] + 1) {
      if (out) out[n] = make_int2(sn_nodes[start], (start - b) | ((i - start) << 16));
      ++n;
      start = i;
    }
  }
  if (counts) counts[s] = n;
}

// per-slice descriptor, 48 bytes = three 16-byte loads, everything the assembling warp needs to start its copies:
//   {matrix base, incidence base, node-list base, w | iw<<6 | nn<<12 | l_first<<22}   (l_first = position of the slice's
//   first row in its node list; the slice's own rows are consecutive there because the list is sorted)
//   {run base, number of runs, run 0}, {run 1, run 2}   (the first three runs inline: a regular numbering has exactly three)
__global__ void k_fill_sdesc(int nslices, const int* __restrict__ slice_ptr, const int* __restrict__ inc_slice_ptr,
                             const int* __restrict__ sn_ptr, const int* __restrict__ sn_nodes,
                             const int* __restrict__ run_ptr, const int2* __restrict__ runs, int4* __restrict__ sdesc) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  const int base = slice_ptr[s], ibase = inc_slice_ptr[s], nbase = sn_ptr[s];
  const unsigned w = (unsigned)(slice_ptr[s + 1] - base) >> 5, iw = (unsigned)(inc_slice_ptr[s + 1] - ibase) >> 5;
  const unsigned nn = (unsigned)(sn_ptr[s + 1] - nbase);
  const unsigned lf = (unsigned)find_slot(sn_nodes + nbase, (int)nn, s * kSlice);
  const int rb = run_ptr[s], nr = run_ptr[s + 1] - rb;
  int2 r[3] = {make_int2(0, 0), make_int2(0, 0), make_int2(0, 0)};
  for (int j = 0; j < 3 && j < nr; ++j) r[j] = runs[rb + j];
  sdesc[3 * (size_t)s + 0] = make_int4(base, ibase, nbase, (int)(w | (iw << 6) | (nn << 12) | (lf << 22)));
  sdesc[3 * (size_t)s + 1] = make_int4(rb, nr, r[0].x, r[0].y);
  sdesc[3 * (size_t)s + 2] = make_int4(r[1].x, r[1].y, r[2].x, r[2].y);
}

__global__ void k_min_deg(const int* __restrict__ ptr, int nrows, int* __restrict__ out_min) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < nrows) atomicMin(out_min, ptr[row + 1] - ptr[row]);
}

// Temporaries come from the device's stream-ordered memory pool (cudaMallocAsync on the library stream, which is the only
// stream that touches them): a second context in the same process otherwise pays hundreds of milliseconds in the
// legacy cudaMalloc/cudaFree path for the same few GB of sort buffers (measured: 29 ms -> 770 ms for this function).
struct Tmp {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit Tmp(cudaStream_t s) : st(s) {}
  ~Tmp() { for (void* p : ptrs) cudaFreeAsync(p, st); }
  template <typename T> cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMallocAsync((void**)p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

inline int bits_for(uint64_t n) { int b = 1; while (b < 63 && (1ull << b) < n) ++b; return b; }

}  // namespace

void free_pattern(Context* ctx) {
  Pattern& P = ctx->pat;
  cudaFree(P.slice_ptr); cudaFree(P.col); cudaFree(P.rowptr); cudaFree(P.csrcol);
  cudaFree(P.inc_slice_ptr); cudaFree(P.inc_tri); cudaFree(P.inc_pack); cudaFree(P.arinc); cudaFree(P.sdesc);
  cudaFree(P.sn_ptr); cudaFree(P.sn_nodes); cudaFree(P.runs); cudaFree(P.rowinfo);
  P = Pattern();
}

int build_pattern(Context* ctx) {
  Pattern& P = ctx->pat;
  free_pattern(ctx);
  cudaStream_t st = ctx->stream;
  const int nme = ctx->nme, n_own = ctx->n_own, n_loc = ctx->n_loc;
  if ((int64_t)nme * 4 + 3 >= (1LL << 31)) { ctx->fail(DDPS_ERR_ARG, "too many triangles per GPU (packed id overflow)"); return DDPS_ERR_ARG; }
  P.nrows = n_own; P.ncols = n_loc; P.nslices = (n_own + kSlice - 1) / kSlice;
  const int nsl = P.nslices;
  Tmp tmp(st);
  const int T = 256;

  // ---- 1. incidence pairs sorted by node (stable -> increasing (triangle,corner) per node, i.e. the
  //         reference's duplicate-summation order of sparse()) -------------------------
  const int64_t n3 = 3LL * nme;
  int *ik_in, *ik_out, *iv_in, *iv_out, *inc_ptr;
  DDPS_CUDA(tmp.alloc(&ik_in, n3)); DDPS_CUDA(tmp.alloc(&ik_out, n3));
  DDPS_CUDA(tmp.alloc(&iv_in, n3)); DDPS_CUDA(tmp.alloc(&iv_out, n3));
  DDPS_CUDA(tmp.alloc(&inc_ptr, (size_t)n_own + 1));
  k_make_inc_pairs<<<(nme + T - 1) / T, T, 0, st>>>(ctx->elem, nme, ik_in, iv_in);
  {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, ik_in, ik_out, iv_in, iv_out, n3, 0, bits_for(n_loc), st);
    void* ws; DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortPairs(ws, bytes, ik_in, ik_out, iv_in, iv_out, n3, 0, bits_for(n_loc), st));
  }
  k_lower_bound<int><<<(n_own + 1 + T - 1) / T, T, 0, st>>>(ik_out, n3, n_own + 1, 0, inc_ptr);

  // ---- 2. adjacency keys (row<<32|col), sort + unique -> CSR with sorted columns ------------------
  const int64_t nk = 6LL * nme + n_own;
  unsigned long long *ak_in, *ak_out, *ak_uni;
  int* d_num;
  DDPS_CUDA(tmp.alloc(&ak_in, nk)); DDPS_CUDA(tmp.alloc(&ak_out, nk)); DDPS_CUDA(tmp.alloc(&d_num, 4));
  {
    int64_t nt = (int64_t)nme + n_own;
    k_make_adj_keys<<<(unsigned)((nt + T - 1) / T), T, 0, st>>>(ctx->elem, nme, n_own, ak_in);
    size_t bytes = 0;
    int endbit = 32 + bits_for(n_loc);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, ak_in, ak_out, nk, 0, endbit, st);
    void* ws; DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortKeys(ws, bytes, ak_in, ak_out, nk, 0, endbit, st));
    ak_uni = ak_in;  // reuse as output of unique
    size_t b2 = 0;
    cub::DeviceSelect::Unique(nullptr, b2, ak_out, ak_uni, d_num, nk, st);
    void* ws2; DDPS_CUDA(tmp.alloc((char**)&ws2, b2));
    DDPS_CUDA(cub::DeviceSelect::Unique(ws2, b2, ak_out, ak_uni, d_num, nk, st));
  }
  int nuniq = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nuniq, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaMalloc(&P.rowptr, ((size_t)n_own + 1) * sizeof(int)));
  k_lower_bound<unsigned long long><<<(n_own + 1 + T - 1) / T, T, 0, st>>>(ak_uni, nuniq, n_own + 1, 32, P.rowptr);
  int nnz = 0;
  DDPS_CUDA(cudaMemcpyAsync(&nnz, P.rowptr + n_own, sizeof(int), cudaMemcpyDeviceToHost, st));
  DDPS_CUDA(cudaStreamSynchronize(st));
  P.nnz = nnz;
  DDPS_CUDA(cudaMalloc(&P.csrcol, std::max<size_t>(nnz, 1) * sizeof(int)));
  k_low32<<<(unsigned)((nnz + T - 1) / T), T, 0, st>>>(ak_uni, nnz, P.csrcol);

  // ---- 3. slice-interleaved pattern ---------------------------------------------------------------
  long long *w32, *w32scan;
  int* d_max;
  DDPS_CUDA(tmp.alloc(&w32, (size_t)nsl + 1)); DDPS_CUDA(tmp.alloc(&w32scan, (size_t)nsl + 1));
  DDPS_CUDA(tmp.alloc(&d_max, 4));
  auto make_slices = [&](const int* ptr, int** slice_ptr_out, int64_t* entries_out, int* maxlen_out) -> int {
    DDPS_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    DDPS_CUDA(cudaMemsetAsync(w32, 0, ((size_t)nsl + 1) * sizeof(long long), st));
    k_slice_width<<<(nsl * 32 + T - 1) / T, T, 0, st>>>(ptr, n_own, nsl, w32, d_max);
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, w32, w32scan, nsl + 1, st);
    void* ws; DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, w32, w32scan, nsl + 1, st));
    long long total = 0;
    DDPS_CUDA(cudaMemcpyAsync(&total, w32scan + nsl, sizeof(long long), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaMemcpyAsync(maxlen_out, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (total >= (1LL << 31)) { ctx->fail(DDPS_ERR_ARG, "pattern too large for 32-bit entry offsets on one GPU"); return DDPS_ERR_ARG; }
    *entries_out = total;
    DDPS_CUDA(cudaMalloc(slice_ptr_out, ((size_t)nsl + 1) * sizeof(int)));
    k_narrow<<<(nsl + 1 + T - 1) / T, T, 0, st>>>(w32scan, nsl + 1, *slice_ptr_out);
    return 0;
  };
  int rc = make_slices(P.rowptr, &P.slice_ptr, &P.nentries, &P.max_row_len);
  if (rc) return rc;
  if (P.max_row_len > kMaxRowLen) {
    ctx->fail(DDPS_ERR_ARG, "a node has more than " + std::to_string(kMaxRowLen - 1) + " neighbours; unsupported mesh");
    return DDPS_ERR_ARG;
  }
  DDPS_CUDA(cudaMalloc(&P.col, std::max<size_t>(P.nentries, 1) * sizeof(int)));
  k_fill_sell_col<<<(nsl * 32 + T - 1) / T, T, 0, st>>>(P.rowptr, P.csrcol, n_own, nsl, P.slice_ptr, P.col);

  // ---- 4. slice-interleaved incidence records with pre-computed slots -----------------------------
  int maxdeg = 0;
  rc = make_slices(inc_ptr, &P.inc_slice_ptr, &P.inc_entries, &maxdeg);
  if (rc) return rc;
  P.max_inc = maxdeg;
  {
    int* d_min = d_max;
    int big = 1 << 30;
    DDPS_CUDA(cudaMemcpyAsync(d_min, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    k_min_deg<<<(n_own + T - 1) / T, T, 0, st>>>(inc_ptr, n_own, d_min);
    int mind = 0;
    DDPS_CUDA(cudaMemcpyAsync(&mind, d_min, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    if (n_own > 0 && mind == 0) {
      ctx->fail(DDPS_ERR_ARG, "mesh has a node that belongs to no triangle (reference: singular matrix in \\)");
      return DDPS_ERR_ARG;
    }
  }
  // ---- 5. per slice (32 rows = one assembling warp): sorted list of the distinct nodes its rows touch ----------
  {
    unsigned long long *bk_in, *bk_out;
    DDPS_CUDA(tmp.alloc(&bk_in, (size_t)nnz)); DDPS_CUDA(tmp.alloc(&bk_out, (size_t)nnz));
    k_make_block_keys<<<(n_own + T - 1) / T, T, 0, st>>>(P.rowptr, P.csrcol, n_own, bk_in);
    size_t bytes = 0;
    const int endbit = 32 + bits_for((uint64_t)nsl + 1);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, bk_in, bk_out, nnz, 0, endbit, st);
    void* ws; DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceRadixSort::SortKeys(ws, bytes, bk_in, bk_out, nnz, 0, endbit, st));
    size_t b2 = 0;
    cub::DeviceSelect::Unique(nullptr, b2, bk_out, bk_in, d_num, nnz, st);
    void* ws2; DDPS_CUDA(tmp.alloc((char**)&ws2, b2));
    DDPS_CUDA(cub::DeviceSelect::Unique(ws2, b2, bk_out, bk_in, d_num, nnz, st));
    int nlist = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nlist, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    DDPS_CUDA(cudaMalloc(&P.sn_ptr, ((size_t)nsl + 1) * sizeof(int)));
    DDPS_CUDA(cudaMalloc(&P.sn_nodes, std::max<size_t>(nlist, 1) * sizeof(int)));
    k_lower_bound<unsigned long long><<<(nsl + 1 + T - 1) / T, T, 0, st>>>(bk_in, nlist, nsl + 1, 32, P.sn_ptr);
    k_low32<<<(unsigned)((nlist + T - 1) / T), T, 0, st>>>(bk_in, nlist, P.sn_nodes);
    DDPS_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
    k_block_maxlen<<<(nsl + T - 1) / T, T, 0, st>>>(P.sn_ptr, nsl, d_max);
    DDPS_CUDA(cudaMemcpyAsync(&P.max_sn, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    P.sn_total = nlist;
    if (P.max_sn > 1023 || P.max_row_len > 32 || P.max_inc > 63) {
      ctx->fail(DDPS_ERR_ARG, "mesh connectivity too dense for the packed incidence records (a node with > 31 neighbours "
                              "or > 63 triangles, or a 32-row slice touching > 1023 nodes)");
      return DDPS_ERR_ARG;
    }
  }
  DDPS_CUDA(cudaMalloc(&P.inc_tri, std::max<size_t>(P.inc_entries, 1) * sizeof(int)));
  DDPS_CUDA(cudaMalloc(&P.inc_pack, std::max<size_t>(P.inc_entries, 1) * sizeof(unsigned)));
  DDPS_CUDA(cudaMalloc(&P.arinc, std::max<size_t>(P.inc_entries, 1) * sizeof(double)));
  DDPS_CUDA(cudaMalloc(&P.sdesc, std::max<size_t>(nsl, 1) * 3 * sizeof(int4)));
  DDPS_CUDA(cudaMalloc(&P.rowinfo, std::max<size_t>(n_own, 1) * sizeof(unsigned short)));
  k_fill_incidence<<<(nsl * 32 + T - 1) / T, T, 0, st>>>(inc_ptr, iv_out, ctx->elem, ctx->xy, P.rowptr, P.csrcol, n_own, nsl,
                                                        P.inc_slice_ptr, P.sn_ptr, P.sn_nodes, P.inc_tri, P.inc_pack,
                                                        P.arinc, P.rowinfo);
  // runs of consecutive nodes in the slice lists (count, scan, write), then the descriptors
  {
    int *cnt, *run_ptr;
    DDPS_CUDA(tmp.alloc(&cnt, (size_t)nsl + 1)); DDPS_CUDA(tmp.alloc(&run_ptr, (size_t)nsl + 1));
    DDPS_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)nsl + 1) * sizeof(int), st));
    k_slice_runs<<<(nsl + T - 1) / T, T, 0, st>>>(nsl, P.sn_ptr, P.sn_nodes, nullptr, cnt, nullptr);
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, run_ptr, nsl + 1, st);
    void* ws; DDPS_CUDA(tmp.alloc((char**)&ws, bytes));
    DDPS_CUDA(cub::DeviceScan::ExclusiveSum(ws, bytes, cnt, run_ptr, nsl + 1, st));
    int nruns = 0;
    DDPS_CUDA(cudaMemcpyAsync(&nruns, run_ptr + nsl, sizeof(int), cudaMemcpyDeviceToHost, st));
    DDPS_CUDA(cudaStreamSynchronize(st));
    P.nruns = nruns;
    DDPS_CUDA(cudaMalloc(&P.runs, std::max<size_t>(nruns, 1) * sizeof(int2)));
    k_slice_runs<<<(nsl + T - 1) / T, T, 0, st>>>(nsl, P.sn_ptr, P.sn_nodes, run_ptr, nullptr, P.runs);
    k_fill_sdesc<<<(nsl + T - 1) / T, T, 0, st>>>(nsl, P.slice_ptr, P.inc_slice_ptr, P.sn_ptr, P.sn_nodes, run_ptr, P.runs,
                                                 P.sdesc);
  }
  DDPS_CUDA(cudaStreamSynchronize(st));
  DDPS_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Recursive coordinate bisection of the nodes into nparts (power of two not required) parts.
// Host code: runs once per mesh; each level splits the current index range at the weighted median of
// the longer bounding-box axis (ties broken by node id so every rank computes the same answer).
// ------------------------------------------------------------------------------------------------
static void rcb(const double* nodes, int64_t* idx, int64_t n, int part0, int nparts, int32_t* owner) {
  if (nparts == 1) {
    for (int64_t i = 0; i < n; ++i) owner[idx[i]] = part0;
    return;
  }
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int64_t i = 0; i < n; ++i) {
    const double* p = nodes + 2 * idx[i];
    for (int d = 0; d < 2; ++d) { lo[d] = std::min(lo[d], p[d]); hi[d] = std::max(hi[d], p[d]); }
  }
  int axis = (hi[1] - lo[1] > hi[0] - lo[0]) ? 1 : 0;
  int pl = nparts / 2;
  int64_t nl = (int64_t)((double)n * pl / nparts + 0.5);
  auto cmp = [&](int64_t a, int64_t b) {
    double ca = nodes[2 * a + axis], cb = nodes[2 * b + axis];
    if (ca != cb) return ca < cb;
    double oa = nodes[2 * a + 1 - axis], ob = nodes[2 * b + 1 - axis];
    if (oa != ob) return oa < ob;
    return a < b;
  };
  std::nth_element(idx, idx + nl, idx + n, cmp);
  rcb(nodes, idx, nl, part0, pl, owner);
  rcb(nodes, idx + nl, n - nl, part0 + pl, nparts - pl, owner);
}

void partition_nodes_rcb(const double* nodes, int64_t nq, int nparts, int32_t* owner) {
  std::vector<int64_t> idx(nq);
  std::iota(idx.begin(), idx.end(), (int64_t)0);
  rcb(nodes, idx.data(), nq, 0, nparts, owner);
}

}  // namespace ddps
