// comm.cu -- the multi-GPU side of the C-ABI: slab partition of the elements over ranks, an NCCL
// communicator the host code can create without touching NCCL itself, and the exchange that completes the rows of the
// sharded global matrix / load vector whose DOFs are touched by elements of more than one rank.
//
// Functional spec: the scatter loop of dune/python/test/poisson.py:113-125 adds every element's K_e into A[gi, gj]; when
// the elements are split over ranks, row gi is complete only after the partial sums of all ranks whose elements contain
// DOF gi have been added.  With slabs along the slowest direction and g_d = s_d - p_d + l_d (dune/iga/nurbsbasis.hh:576)
// the rows a rank touches are ONE contiguous range (whole DOF layers), so the rows two ranks share are the intersection
// of two ranges -- not only direct neighbours: a slab thinner than the degree shares DOF layers with ranks further away.
// The hook on the reference side is the grid view's communicate() (dune/iga/nurbsleafgridview.hh:163-165).
//
// NCCL is resolved with dlsym at run time (no link-time dependency); point-to-point ncclSend / ncclRecv inside one
// group on the caller's stream, then one kernel per contiguous piece adds the contributions IN ASCENDING RANK ORDER, so
// that every rank holding a shared row ends up with bit-identical values.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "igab_internal.cuh"

namespace igab {
namespace {

struct NcclUid {
  char internal[128];
};
struct NcclApi {
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommCount)(const void*, int*) = nullptr;
  int (*CommUserRank)(const void*, int*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

const NcclApi& nccl() {
  static const NcclApi api = [] {
    NcclApi a;
    void* h = nullptr;
    auto sym = [&](const char* name) -> void* {
      void* s = dlsym(RTLD_DEFAULT, name);
      if (!s) {
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h) s = dlsym(h, name);
      }
      return s;
    };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.CommCount = (decltype(a.CommCount))sym("ncclCommCount");
    a.CommUserRank = (decltype(a.CommUserRank))sym("ncclCommUserRank");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommCount && a.CommUserRank && a.GroupStart &&
           a.GroupEnd && a.Send && a.Recv;
    return a;
  }();
  return api;
}

igab_status nccl_fail(int rc, const char* what) {
  const NcclApi& a = nccl();
  set_error(std::string("NCCL error in ") + what + ": " +
            (a.GetErrorString ? a.GetErrorString(rc) : ("code " + std::to_string(rc)).c_str()));
  return IGAB_NCCL_ERROR;
}
igab_status need_nccl() {
  if (!nccl().ok) {
    set_error("libnccl.so.2 could not be resolved in this process");
    return IGAB_NCCL_ERROR;
  }
  return IGAB_OK;
}

constexpr int kMaxContrib = 32;  // ranks (this one included) that may share one row

// Contributions to a contiguous piece [begin, end) of an array: `own` points at index `begin` of this rank's partial
// sums; contributor k (ascending rank; slot `self` is this rank) covers [cb[k], ce[k]) and, unless it is this rank, its
// values sit in buf[k][i - cb[k]].
struct AddDesc {
  int ncontrib, self;
  long long begin, end;
  long long cb[kMaxContrib], ce[kMaxContrib];
  const double* buf[kMaxContrib];
};

__global__ void __launch_bounds__(256) ordered_add_kernel(AddDesc d, double* __restrict__ own) {
  const long long i = d.begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.end) return;
  double acc = 0.0;
  bool first = true, shared = false;
  for (int k = 0; k < d.ncontrib; ++k) {
    if (i < d.cb[k] || i >= d.ce[k]) continue;
    const double v = (k == d.self) ? own[i - d.begin] : d.buf[k][i - d.cb[k]];
    if (k != d.self) shared = true;
    acc = first ? v : acc + v;
    first = false;
  }
  if (shared) own[i - d.begin] = acc;
}

// first-DOF layer (span - degree) of every element layer of one direction, exactly as igab_patch_create builds it
bool feq8(double a, double b) {
  const double eps = 8.0 * 2.220446049250313e-16;
  return std::fabs(a - b) <= eps * std::max(std::fabs(a), std::fabs(b));
}
std::vector<int> first_dof_layers(int p, const double* U, int n) {
  std::vector<double> uq;
  for (int i = 0; i < n; ++i)
    if (uq.empty() || !feq8(uq.back(), U[i])) uq.push_back(U[i]);
  std::vector<int> first;
  for (size_t e = 0; e + 1 < uq.size(); ++e) {
    int s;
    if (uq[e] <= U[0]) s = p;
    else if (uq[e] >= U[n - 1]) s = n - p - 2;
    else s = (int)(std::upper_bound(U + p - 1 + e, U + n, uq[e]) - U) - 1;
    first.push_back(s - p);
  }
  return first;
}

// touched row range of every rank (untrimmed numbering) from the per-direction sizes and the last direction's layers
igab_status plan_rows(int dim, const int* ncp, const int* nel, const std::vector<int>& firstLast, int pLast, int ncomp,
                      int world, const int64_t* elem_bounds, int64_t* row_begin, int64_t* row_end) {
  long long perLayer = 1, dofStride = 1;
  for (int d = 0; d + 1 < dim; ++d) {
    perLayer *= nel[d];
    dofStride *= ncp[d];
  }
  const int layers = nel[dim - 1];
  for (int r = 0; r < world; ++r) {
    long long eb, ee;
    if (elem_bounds) {
      eb = elem_bounds[r];
      ee = elem_bounds[r + 1];
      if (eb < 0 || ee < eb || ee > perLayer * layers || eb % perLayer || ee % perLayer ||
          (r == 0 && eb != 0) || (r == world - 1 && ee != perLayer * layers)) {
        set_error("interface rows: element bounds must be ascending multiples of the layer size covering the patch "
                  "(slabs of whole layers of the slowest direction)");
        return IGAB_UNSUPPORTED;
      }
    } else {
      const long long base = layers / world, rem = layers % world;
      const long long lo = r * base + std::min<long long>(r, rem);
      eb = lo * perLayer;
      ee = (lo + base + (r < rem ? 1 : 0)) * perLayer;
    }
    if (ee == eb) {
      row_begin[r] = row_end[r] = 0;
      continue;
    }
    const int l0 = (int)(eb / perLayer), l1 = (int)(ee / perLayer) - 1;
    row_begin[r] = (long long)firstLast[l0] * dofStride * ncomp;
    row_end[r] = (long long)(firstLast[l1] + pLast + 1) * dofStride * ncomp;
  }
  return IGAB_OK;
}

}  // namespace
}  // namespace igab

using namespace igab;

extern "C" {

igab_status igab_partition_slab(int dim, const int32_t* n_elem_dir, int world, int rank, int64_t* elem_begin,
                                int64_t* elem_end) {
  if (dim < 1 || dim > kMaxDim || !n_elem_dir || world < 1 || rank < 0 || rank >= world || !elem_begin || !elem_end) {
    set_error("partition_slab: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  long long perLayer = 1;
  for (int d = 0; d + 1 < dim; ++d) perLayer *= n_elem_dir[d];
  const long long layers = n_elem_dir[dim - 1], base = layers / world, rem = layers % world;
  const long long lo = rank * base + std::min<long long>(rank, rem);
  *elem_begin = lo * perLayer;
  *elem_end = (lo + base + (rank < rem ? 1 : 0)) * perLayer;
  return IGAB_OK;
}

igab_status igab_partition_weighted(int64_t n_elem, const int64_t* counts, int world, int64_t per_layer,
                                    int64_t* bounds) {
  if (n_elem < 0 || (n_elem > 0 && !counts) || world < 1 || per_layer < 1 || !bounds) {
    set_error("partition_weighted: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  long long total = 0;
  for (int64_t e = 0; e < n_elem; ++e) {
    if (counts[e] < 0) {
      set_error("partition_weighted: negative point count");
      return IGAB_INVALID_ARGUMENT;
    }
    total += counts[e];
  }
  bounds[0] = 0;
  long long run = 0;
  int64_t e = 0;
  for (int r = 1; r < world; ++r) {
    // first element index b with (points of elements < b) >= total * r / world, compared without division
    while (e < n_elem && (__int128)run * world < (__int128)total * r) run += counts[e++];
    int64_t b = e;
    if (per_layer > 1) b = (b + per_layer / 2) / per_layer * per_layer;
    bounds[r] = std::min<int64_t>(std::max<int64_t>(b, bounds[r - 1]), n_elem);
  }
  bounds[world] = n_elem;
  return IGAB_OK;
}

igab_status igab_interface_row_ranges(int dim, const int* degree, const int* nknots, const double* const* knots,
                                      int ncomp, int world, const int64_t* elem_bounds, int64_t* row_begin,
                                      int64_t* row_end) {
  if (dim < 1 || dim > kMaxDim || !degree || !nknots || !knots || ncomp < 1 || world < 1 || !row_begin || !row_end) {
    set_error("interface_row_ranges: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  int ncp[kMaxDim], nel[kMaxDim];
  std::vector<int> firstLast;
  for (int d = 0; d < dim; ++d) {
    if (degree[d] < 1 || nknots[d] < 2 * (degree[d] + 1) || !knots[d]) {
      set_error("interface_row_ranges: invalid knot vector");
      return IGAB_INVALID_ARGUMENT;
    }
    ncp[d] = nknots[d] - degree[d] - 1;
    std::vector<int> f = first_dof_layers(degree[d], knots[d], nknots[d]);
    nel[d] = (int)f.size();
    if (d == dim - 1) firstLast = f;
  }
  return plan_rows(dim, ncp, nel, firstLast, degree[dim - 1], ncomp, world, elem_bounds, row_begin, row_end);
}

igab_status igab_comm_unique_id(void* id128) {
  if (!id128) return IGAB_INVALID_ARGUMENT;
  if (igab_status st = need_nccl()) return st;
  NcclUid id;
  const int rc = nccl().GetUniqueId(&id);
  if (rc) return nccl_fail(rc, "ncclGetUniqueId");
  std::memcpy(id128, &id, sizeof(id));
  return IGAB_OK;
}

igab_status igab_comm_create(int rank, int world, const void* id128, void** comm) {
  if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) {
    set_error("comm_create: invalid argument");
    return IGAB_INVALID_ARGUMENT;
  }
  if (igab_status st = need_nccl()) return st;
  NcclUid id;
  std::memcpy(&id, id128, sizeof(id));
  *comm = nullptr;
  const int rc = nccl().CommInitRank(comm, world, id, rank);
  if (rc) return nccl_fail(rc, "ncclCommInitRank");
  return IGAB_OK;
}

igab_status igab_comm_destroy(void* comm) {
  if (!comm) return IGAB_OK;
  if (igab_status st = need_nccl()) return st;
  const int rc = nccl().CommDestroy(comm);
  if (rc) return nccl_fail(rc, "ncclCommDestroy");
  return IGAB_OK;
}

igab_status igab_comm_info(void* comm, int* rank, int* world) {
  if (!comm) return IGAB_INVALID_ARGUMENT;
  if (igab_status st = need_nccl()) return st;
  int rc = 0;
  if (world && (rc = nccl().CommCount(comm, world))) return nccl_fail(rc, "ncclCommCount");
  if (rank && (rc = nccl().CommUserRank(comm, rank))) return nccl_fail(rc, "ncclCommUserRank");
  return IGAB_OK;
}

igab_status igab_exchange_interface_rows(igab_patch* p, int ncomp, const int64_t* elem_bounds, double* values, double* b,
                                         igab_memspace space, void* nccl_comm, void* stream_) {
  if (!p || !nccl_comm || (!values && !b)) {
    set_error("exchange_interface_rows: null argument (needs a communicator and values and / or b)");
    return IGAB_INVALID_ARGUMENT;
  }
  igab_status st = need_nccl();
  if (st) return st;
  const NcclApi& N = nccl();
  int rank = 0, world = 1, rc = 0;
  if ((rc = N.CommCount(nccl_comm, &world))) return nccl_fail(rc, "ncclCommCount");
  if ((rc = N.CommUserRank(nccl_comm, &rank))) return nccl_fail(rc, "ncclCommUserRank");
  cudaStream_t stream = (cudaStream_t)stream_;
  const PatchDev& D = p->dev;
  const int dl = D.dim - 1;
  std::vector<int> firstLast(D.nel[dl]);
  for (int e = 0; e < D.nel[dl]; ++e) firstLast[e] = p->hspan[dl][e] - D.degree[dl];
  std::vector<int64_t> rb(world), re(world);
  if ((st = plan_rows(D.dim, D.ncp, D.nel, firstLast, D.degree[dl], ncomp, world, elem_bounds, rb.data(), re.data())))
    return st;
  int64_t nrows = 0, nnz = 0;
  if ((st = csr_pattern(p, ncomp, &nrows, &nnz, nullptr, nullptr, IGAB_HOST, stream))) return st;
  if (p->d_trim) {
    // compacted numbering (prepareForTrim, nurbsbasis.hh:599-615): the map is monotone, so a range of original rows is a
    // range of compacted rows; rows of DOFs no live element touches simply drop out of it
    if (p->dofc_map_host.empty()) {
      p->dofc_map_host.resize(p->ndof_untrimmed);
      IGAB_CUDA_TRY(cudaMemcpy(p->dofc_map_host.data(), p->d_dofc_map, sizeof(int32_t) * p->ndof_untrimmed,
                               cudaMemcpyDeviceToHost));
      // lower[g] = compacted index of the first kept DOF >= g
      long long next = p->ndof_compact;
      for (long long g = p->ndof_untrimmed - 1; g >= 0; --g) {
        if (p->dofc_map_host[g] >= 0) next = p->dofc_map_host[g];
        p->dofc_map_host[g] = (int32_t)next;
      }
    }
    auto lower = [&](int64_t row) {
      const int64_t g = row / ncomp;
      return (int64_t)(g >= p->ndof_untrimmed ? p->ndof_compact : p->dofc_map_host[g]) * ncomp;
    };
    for (int r = 0; r < world; ++r) {
      rb[r] = lower(rb[r]);
      re[r] = lower(re[r]);
    }
  }
  // peers and the (at most two) disjoint pieces of this rank's rows that somebody else also touches
  struct Peer {
    int rank;
    int64_t r0, r1;
  };
  std::vector<Peer> peers;
  for (int s = 0; s < world; ++s) {
    if (s == rank) continue;
    const int64_t r0 = std::max(rb[s], rb[rank]), r1 = std::min(re[s], re[rank]);
    if (r1 > r0) peers.push_back({s, r0, r1});
  }
  if (peers.empty()) return IGAB_OK;
  if ((int)peers.size() + 1 > kMaxContrib) {
    set_error("exchange_interface_rows: more than 31 ranks share a row (slabs far thinner than the degree)");
    return IGAB_UNSUPPORTED;
  }
  std::vector<std::pair<int64_t, int64_t>> pieces;
  {
    int64_t loEnd = rb[rank], hiBeg = re[rank];
    for (const Peer& q : peers) {
      if (q.rank < rank) loEnd = std::max(loEnd, q.r1);
      else hiBeg = std::min(hiBeg, q.r0);
    }
    if (loEnd >= hiBeg) pieces.push_back({rb[rank], re[rank]});
    else {
      if (loEnd > rb[rank]) pieces.push_back({rb[rank], loEnd});
      if (hiBeg < re[rank]) pieces.push_back({hiBeg, re[rank]});
    }
  }
  const std::vector<long long>& rowptr = p->csr_rowptr_host;
  for (int pass = 0; pass < 2; ++pass) {
    double* arr = pass == 0 ? values : b;
    if (!arr) continue;
    auto idx = [&](int64_t row) { return pass == 0 ? (int64_t)rowptr[row] : row; };
    // scratch: [receive buffers of all peers | staged own pieces (HOST space)]
    size_t need = 0;
    for (const Peer& q : peers) need += (size_t)(idx(q.r1) - idx(q.r0));
    const size_t recvTotal = need;
    if (space == IGAB_HOST)
      for (auto& pc : pieces) need += (size_t)(idx(pc.second) - idx(pc.first));
    if (p->xch_cap < need) {
      IGAB_CUDA_TRY(cudaStreamSynchronize(stream));
      if (p->xch_buf) cudaFree(p->xch_buf);
      p->xch_buf = nullptr;
      p->xch_cap = 0;
      IGAB_CUDA_TRY(cudaMalloc(&p->xch_buf, sizeof(double) * std::max<size_t>(need, 1)));
      p->xch_cap = need;
    }
    std::vector<double*> own(pieces.size());
    {
      double* cur = p->xch_buf + recvTotal;
      for (size_t k = 0; k < pieces.size(); ++k) {
        const int64_t a = idx(pieces[k].first), z = idx(pieces[k].second);
        if (space == IGAB_HOST) {
          own[k] = cur;
          IGAB_CUDA_TRY(cudaMemcpyAsync(cur, arr + a, sizeof(double) * (z - a), cudaMemcpyHostToDevice, stream));
          cur += z - a;
        } else {
          own[k] = arr + a;
        }
      }
    }
    auto piece_of = [&](int64_t r0) {
      for (size_t k = 0; k < pieces.size(); ++k)
        if (r0 >= pieces[k].first && r0 < pieces[k].second) return k;
      return (size_t)0;
    };
    std::vector<double*> recv(peers.size());
    if ((rc = N.GroupStart())) return nccl_fail(rc, "ncclGroupStart");
    {
      double* cur = p->xch_buf;
      for (size_t k = 0; k < peers.size() && !rc; ++k) {
        const Peer& q = peers[k];
        const int64_t a = idx(q.r0), z = idx(q.r1);
        const size_t pk = piece_of(q.r0);
        recv[k] = cur;
        cur += z - a;
        rc = N.Send(own[pk] + (a - idx(pieces[pk].first)), (size_t)(z - a), 8 /*ncclDouble*/, q.rank, nccl_comm, stream);
        if (!rc) rc = N.Recv(recv[k], (size_t)(z - a), 8, q.rank, nccl_comm, stream);
      }
    }
    const int rcEnd = N.GroupEnd();
    if (rc) return nccl_fail(rc, "ncclSend / ncclRecv");
    if (rcEnd) return nccl_fail(rcEnd, "ncclGroupEnd");
    for (size_t k = 0; k < pieces.size(); ++k) {
      AddDesc d{};
      d.begin = idx(pieces[k].first);
      d.end = idx(pieces[k].second);
      d.self = -1;
      for (size_t j = 0; j <= peers.size(); ++j) {
        // ascending rank: this rank's own partial sums take their place among the peers'
        const bool selfHere = d.self < 0 && (j == peers.size() || peers[j].rank > rank);
        if (selfHere) {
          d.self = d.ncontrib;
          d.cb[d.ncontrib] = d.begin;
          d.ce[d.ncontrib] = d.end;
          d.buf[d.ncontrib] = nullptr;
          ++d.ncontrib;
        }
        if (j == peers.size()) break;
        const int64_t a = std::max<int64_t>(idx(peers[j].r0), d.begin), z = std::min<int64_t>(idx(peers[j].r1), d.end);
        if (z <= a) continue;
        d.cb[d.ncontrib] = idx(peers[j].r0);
        d.ce[d.ncontrib] = idx(peers[j].r1);
        d.buf[d.ncontrib] = recv[j];
        ++d.ncontrib;
      }
      const long long len = d.end - d.begin;
      if (len <= 0) continue;
      ordered_add_kernel<<<(unsigned)((len + 255) / 256), 256, 0, stream>>>(d, own[k]);
      count_launch();
      IGAB_CUDA_TRY(cudaGetLastError());
      if (space == IGAB_HOST)
        IGAB_CUDA_TRY(cudaMemcpyAsync(arr + d.begin, own[k], sizeof(double) * len, cudaMemcpyDeviceToHost, stream));
    }
    if (space == IGAB_HOST) IGAB_CUDA_TRY(cudaStreamSynchronize(stream));
  }
  return IGAB_OK;
}

}  // extern "C"
