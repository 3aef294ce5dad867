// igab200_adaptor.hh -- header-only C++17 host adaptor that keeps the reference's METHOD NAMES, argument meaning and
// error behaviour for the hot path, and forwards to the batched C-ABI (include/igab200.h) instead of evaluating one
// point at a time on the CPU.  It is what NURBSGrid / NurbsPreBasis host code (which stays C++) would hold in place of
// its per-point members; INTEGRATION.md shows the three-line changes in the reference that bind it.
//
//   reference type / method (file:line)                                   here
//   NURBSPatchData                      dune/iga/nurbspatchdata.hh:17-36   igab::PatchData  (same three members)
//   NurbsLocalBasis::evaluateFunction   dune/iga/nurbsbasis.hh:77-82       LocalBasis::evaluateFunction
//   NurbsLocalBasis::evaluateJacobian   dune/iga/nurbsbasis.hh:87-96       LocalBasis::evaluateJacobian
//   NurbsLocalBasis::partial            dune/iga/nurbsbasis.hh:99-108      LocalBasis::partial
//   NurbsLocalBasis::order / size       dune/iga/nurbsbasis.hh:117-123     LocalBasis::order / size
//   NurbsLocalFiniteElement::bind       dune/iga/nurbsbasis.hh:355-375     LocalView::bind(elementIndex)
//   NurbsPreBasis::indices / size       dune/iga/nurbsbasis.hh:590-598,625 LocalView::index(i), PreBasis::size/dimension
//   NurbsPreBasis::size / dimension / maxNodeSize / sizePerDirection   nurbsbasis.hh:545-633  PreBasis (same names)
//   NurbsLocalCoefficients::localKey    dune/iga/nurbsbasis.hh:150-290     LocalView::localCoefficients().localKey(i)
//   NURBSGeometry::global               dune/iga/nurbsgeometry.hh:133-138  Geometry::global
//   NURBSGeometry::jacobianTransposed   dune/iga/nurbsgeometry.hh:182-196  Geometry::jacobianTransposed
//   NURBSGeometry::integrationElement   dune/iga/nurbsgeometry.hh:200-203  Geometry::integrationElement
//   NURBSGeometry::jacobianInverseTransposed  nurbsgeometry.hh:205-210     Geometry::jacobianInverseTransposed
//   NURBSGeometry::volume               dune/iga/nurbsgeometry.hh:96-113   Geometry::volume
//   NURBSGeometry::local                dune/iga/nurbsgeometry.hh:145-176  Geometry::local
//   NURBSGeometry::secondDerivativeOfPosition / zeroFirstAndSecond...  :257-354  same names
//   NURBSGeometry::normal/unitNormal/metric/secondFundamentalForm/gaussianCurvature  :216-255  same names
//   NURBSGridEntity::getQuadratureRule  dune/iga/nurbsgridentity.hh:120-141 Patch::setQuadratureRule (rule is an input)
//
// How it works: Patch::setQuadratureRule(rule) fixes the rule and a WINDOW of element blocks in page-locked host memory
// (512 MB by default, whatever the patch size).  bind(e) + evaluate*(xi) return the cached row when xi is a point of the
// rule (exact coordinate match -- the quadrature loop of dune/python/test/poisson.py:47-79 or
// dune/iga/test/gridtests.cpp:338-360); the block of elements that holds e is evaluated on the GPU (igab_eval_tensor) the
// first time one of its elements is touched, and the block after it is kept in flight (igab_eval_tensor_begin) while the
// caller works on this one.  Arguments off the rule fall back to a single-point GPU call (igab_eval_points); corners and
// centres are evaluated for a whole block at once.  There is no CPU evaluation path.
//
// Generic over the vector types: anything with operator[] works (Dune::FieldVector / FieldMatrix when DUNE is there,
// std::array in the tests), because DUNE itself is not available in this build environment.
// Error behaviour: the reference asserts / throws (DUNE_THROW, std::logic_error); the adaptor throws igab::Error
// carrying the igab_status and igab_last_error_string().
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <cmath>
#include <utility>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/igab200.h"

namespace igab {

struct Error : std::runtime_error {
  igab_status status;
  Error(igab_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
inline void check(igab_status s) {
  if (s != IGAB_OK) throw Error(s, igab_last_error_string());
}

// Same members as Dune::IGA::NURBSPatchData; control points flat, first direction fastest, rows (x.., w).
template <int dim, int dimworld>
struct PatchData {
  std::array<std::vector<double>, dim> knotSpans;
  std::vector<double> controlPoints;  // [n][dimworld+1]
  std::array<int, dim> degree;
};

template <int dim, int dimworld>
class Patch : public std::enable_shared_from_this<Patch<dim, dimworld>> {
public:
  static constexpr int nh = dim * (dim + 1) / 2;

  explicit Patch(const PatchData<dim, dimworld>& d, const std::vector<uint8_t>* trimFlags = nullptr) : data_(d) {
    std::array<int, dim> nk;
    std::array<const double*, dim> kp;
    for (int i = 0; i < dim; ++i) {
      nk[i] = (int)d.knotSpans[i].size();
      kp[i] = d.knotSpans[i].data();
    }
    check(igab_patch_create(dim, dimworld, d.degree.data(), nk.data(), kp.data(), d.controlPoints.data(),
                            trimFlags ? trimFlags->data() : nullptr, &h_));
    int32_t nb = 0;
    check(igab_patch_sizes(h_, &nElements_, nElDir_.data(), &nb, nullptr));
    nb_ = nb;
    dof_.resize((size_t)nElements_ * nb_);
    check(igab_patch_dofmap(h_, dof_.data(), &nDofs_, IGAB_HOST, nullptr));
  }
  Patch(const Patch&) = delete;
  Patch& operator=(const Patch&) = delete;
  ~Patch() {
    releaseWindow();
    igab_patch_destroy(h_);
  }

  igab_patch* handle() const { return h_; }
  int64_t size0() const { return nElements_; }  // number of elements (codim 0)
  int localSize() const { return nb_; }
  int64_t numDofs() const { return nDofs_; }
  const PatchData<dim, dimworld>& patchData() const { return data_; }
  const int32_t* dofsOf(int64_t e) const { return dof_.data() + (size_t)e * nb_; }

  // NURBSPatch::getDirectIndex<0> / getRealIndex<0> (nurbspatch.hh:599-631): real index = position among the non-empty
  // elements.  getRealIndex throws for an empty element, as the reference does ("No corresponding realIndex").
  int64_t getDirectIndex(int64_t realIndex) const {
    ensureIndexMaps();
    if (realIndex < 0 || realIndex >= (int64_t)directOfReal_.size()) throw Error(IGAB_INVALID_ARGUMENT, "no such real index");
    return directOfReal_[(size_t)realIndex];
  }
  int64_t getRealIndex(int64_t directIndex) const {
    ensureIndexMaps();
    if (directIndex < 0 || directIndex >= nElements_) throw Error(IGAB_INVALID_ARGUMENT, "no such element");
    const int32_t r = realOfDirect_[(size_t)directIndex];
    if (r < 0) throw std::runtime_error("No corresponding realIndex");
    return r;
  }
  int64_t numRealElements() const {
    ensureIndexMaps();
    return (int64_t)directOfReal_.size();
  }

  // The rule every untrimmed element uses (nurbsgridentity.hh:120-133): abscissae / weights per direction on [0,1].
  // derivOrder 1: N, dN, x, Jt, detJ, JinvT are cached; 2 adds d2N, H.
  //
  // The results are NOT held for the whole patch (config C2 would need 289 GB): the cache is a WINDOW of `slots` blocks
  // of consecutive elements in page-locked host memory, windowBytes in total.  Touching an element of a block that is
  // not resident evaluates that block on the GPU (igab_eval_tensor) into the least recently used slot; walking the
  // elements in index order -- the element iteration of the reference, nurbsleafgridview.hh:147-155 -- additionally
  // keeps the NEXT block in flight (igab_eval_tensor_begin), so the caller's per-element work on block b overlaps the
  // kernel and the transfer of block b + 1.  Host memory is bounded by windowBytes whatever the patch size.
  void setQuadratureRule(const std::array<std::vector<double>, dim>& xi, const std::array<std::vector<double>, dim>& w,
                         int derivOrder = 1, size_t windowBytes = size_t(512) << 20, int slots = 3) {
    releaseWindow();
    xi_ = xi;
    w_ = w;
    order_ = derivOrder;
    nq_ = 1;
    for (int i = 0; i < dim; ++i) {
      nq1_[i] = (int)xi[i].size();
      xp_[i] = xi_[i].data();
      nq_ *= nq1_[i];
    }
    // doubles per point of every cached field, in slot order: N dN x Jt detJ JinvT [d2N H]
    const size_t per[8] = {(size_t)nb_, (size_t)nb_ * dim, (size_t)dimworld, (size_t)dim * dimworld, 1,
                           (size_t)dim * dimworld, derivOrder >= 2 ? (size_t)nb_ * nh : 0,
                           derivOrder >= 2 ? (size_t)nh * dimworld : 0};
    size_t perPoint = 0;
    for (size_t v : per) perPoint += v;
    const size_t perElem = perPoint * (size_t)nq_ * sizeof(double);
    if (slots < 2) slots = 2;
    blockElems_ = (int64_t)(windowBytes / (size_t)slots / perElem);
    if (blockElems_ < 1) blockElems_ = 1;
    if (blockElems_ > nElements_) blockElems_ = nElements_;
    numBlocks_ = (nElements_ + blockElems_ - 1) / blockElems_;
    if ((int64_t)slots > numBlocks_) slots = (int)numBlocks_;
    const size_t npts = (size_t)blockElems_ * nq_;
    slotDoubles_ = 0;
    for (int f = 0; f < 8; ++f) {
      fieldOff_[f] = slotDoubles_;
      slotDoubles_ += per[f] * npts;
    }
    slots_.assign((size_t)slots, Slot{});
    for (auto& sl : slots_) {
      void* mem = nullptr;
      check(igab_host_alloc(&mem, slotDoubles_ * sizeof(double)));
      sl.mem = static_cast<double*>(mem);
    }
    haveRule_ = true;
    clock_ = 0;
    blockEvaluations_ = 0;
  }
  bool hasRule() const { return haveRule_; }
  int64_t pointsPerElement() const { return nq_; }
  // what the window holds: host bytes pinned, elements per block, blocks evaluated so far (re-evaluations included)
  size_t windowBytes() const { return slots_.size() * slotDoubles_ * sizeof(double); }
  int64_t blockElements() const { return blockElems_; }
  int64_t blockEvaluations() const { return blockEvaluations_; }
  // weight of rule point q (product of the 1-D weights, first direction fastest)
  double weight(int64_t q) const {
    double r = 1;
    for (int i = 0; i < dim; ++i) {
      r *= w_[i][q % (int64_t)w_[i].size()];
      q /= (int64_t)w_[i].size();
    }
    return r;
  }
  template <class Local>
  void position(int64_t q, Local& xi) const {
    for (int i = 0; i < dim; ++i) {
      xi[i] = xi_[i][q % (int64_t)xi_[i].size()];
      q /= (int64_t)xi_[i].size();
    }
  }

  // index of the rule point equal to `in`, or -1
  template <class Local>
  int64_t findRulePoint(const Local& in) const {
    if (!haveRule_) return -1;
    int64_t q = 0, stride = 1;
    for (int i = 0; i < dim; ++i) {
      int64_t hit = -1;
      for (size_t k = 0; k < xi_[i].size(); ++k)
        if (xi_[i][k] == in[i]) hit = (int64_t)k;
      if (hit < 0) return -1;
      q += hit * stride;
      stride *= (int64_t)xi_[i].size();
    }
    return q;
  }

  // cached rows of element e (the block that holds e is made resident first)
  const double* N(int64_t e, int64_t q) const { return field(0, e, q, (size_t)nb_); }
  const double* dN(int64_t e, int64_t q) const { return field(1, e, q, (size_t)nb_ * dim); }
  const double* x(int64_t e, int64_t q) const { return field(2, e, q, (size_t)dimworld); }
  const double* Jt(int64_t e, int64_t q) const { return field(3, e, q, (size_t)dim * dimworld); }
  double detJ(int64_t e, int64_t q) const { return *field(4, e, q, 1); }
  const double* JinvT(int64_t e, int64_t q) const { return field(5, e, q, (size_t)dim * dimworld); }
  const double* d2N(int64_t e, int64_t q) const { return field(6, e, q, (size_t)nb_ * nh); }
  const double* H(int64_t e, int64_t q) const { return field(7, e, q, (size_t)nh * dimworld); }
  int derivOrder() const { return order_; }

  // corners (k < 2^dim) and centre (k == 2^dim) of element e: evaluated for the WHOLE block of e in one batched call the
  // first time any of them is asked for (the reference evaluates them per call, nurbsgeometry.hh:115-131)
  const double* cornerOrCenter(int64_t e, int k) const {
    if (!haveRule_) return nullptr;
    Slot& sl = slotOf(e);
    constexpr int nc = (1 << dim) + 1;
    if (sl.corners.empty()) {
      const int64_t e0 = sl.block * blockElems_, n = std::min(blockElems_, nElements_ - e0);
      std::vector<int32_t> el((size_t)n * nc);
      std::vector<double> xi((size_t)n * nc * dim);
      for (int64_t a = 0; a < n; ++a)
        for (int c = 0; c < nc; ++c) {
          el[(size_t)a * nc + c] = (int32_t)(e0 + a);
          for (int d = 0; d < dim; ++d) xi[((size_t)a * nc + c) * dim + d] = c == nc - 1 ? 0.5 : ((c >> d) & 1 ? 1.0 : 0.0);
        }
      sl.corners.resize((size_t)n * nc * dimworld);
      igab_eval_out o{};
      o.x = sl.corners.data();
      check(igab_eval_points(h_, n * nc, el.data(), xi.data(), 0, &o, IGAB_HOST, nullptr));
    }
    return sl.corners.data() + ((size_t)(e - sl.block * blockElems_) * nc + k) * dimworld;
  }

  // single point on the GPU (off-rule arguments)
  template <class Local>
  void evalPoint(int64_t e, const Local& in, int order, igab_eval_out& out) const {
    const int32_t el = (int32_t)e;
    double xi[dim];
    for (int i = 0; i < dim; ++i) xi[i] = in[i];
    check(igab_eval_points(h_, 1, &el, xi, order, &out, IGAB_HOST, nullptr));
  }

  // sum over elements of NURBSGeometry::volume() as in gridtests.cpp:338-341 (device reduction)
  double volume() const {
    std::array<int, dim> nq1;
    std::array<const double*, dim> xp, wp;
    for (int i = 0; i < dim; ++i) {
      nq1[i] = (int)xi_[i].size();
      xp[i] = xi_[i].data();
      wp[i] = w_[i].data();
    }
    double r = 0;
    check(igab_reduce_volume(h_, 0, nElements_, nq1.data(), xp.data(), wp.data(), &r, nullptr, nullptr));
    return r;
  }
  // the same on a trimmed patch: full elements with the tensor rule, trimmed elements with their own points
  // (elem [npts], xi [npts][dim] element-local, w [npts]; NURBSGridEntity::getQuadratureRule, fillquadraturerule.hh:8-19)
  double volume(const std::vector<int32_t>& elem, const std::vector<double>& xi, const std::vector<double>& w) const {
    std::array<int, dim> nq1;
    std::array<const double*, dim> xp, wp;
    for (int i = 0; i < dim; ++i) {
      nq1[i] = (int)xi_[i].size();
      xp[i] = xi_[i].data();
      wp[i] = w_[i].data();
    }
    double r = 0;
    check(igab_reduce_volume_trimmed(h_, 0, nElements_, nq1.data(), xp.data(), wp.data(), (int64_t)elem.size(), elem.data(),
                                     xi.data(), w.data(), &r, nullptr, nullptr));
    return r;
  }

private:
  void ensureIndexMaps() const {
    if (!realOfDirect_.empty() || nElements_ == 0) return;
    realOfDirect_.resize((size_t)nElements_);
    directOfReal_.resize((size_t)nElements_);
    int64_t n = 0;
    check(igab_patch_real_index_maps(h_, directOfReal_.data(), realOfDirect_.data(), &n, IGAB_HOST, nullptr));
    directOfReal_.resize((size_t)n);
  }
  mutable std::vector<int32_t> realOfDirect_, directOfReal_;
  struct Slot {
    double* mem = nullptr;        // pinned host memory, all fields of one block back to back
    int64_t block = -1;           // which block it holds (or is being filled with)
    bool inFlight = false;        // igab_eval_tensor_begin queued, not yet waited for
    uint64_t stamp = 0;           // last use (LRU)
    std::vector<double> corners;  // [n][2^dim + 1][dimworld], filled on demand
  };
  igab_eval_out outOf(const Slot& sl) const {
    igab_eval_out o{};
    o.N = sl.mem + fieldOff_[0];
    o.dN = sl.mem + fieldOff_[1];
    o.x = sl.mem + fieldOff_[2];
    o.Jt = sl.mem + fieldOff_[3];
    o.detJ = sl.mem + fieldOff_[4];
    o.JinvT = sl.mem + fieldOff_[5];
    if (order_ >= 2) {
      o.d2N = sl.mem + fieldOff_[6];
      o.H = sl.mem + fieldOff_[7];
    }
    return o;
  }
  Slot* findSlot(int64_t block) const {
    for (auto& sl : slots_)
      if (sl.block == block) return &sl;
    return nullptr;
  }
  Slot& victim(int64_t keep) const {
    Slot* v = nullptr;
    for (auto& sl : slots_)
      if ((keep < 0 || sl.block != keep) && (!v || sl.stamp < v->stamp)) v = &sl;
    return *v;
  }
  void launch(Slot& sl, int64_t block, bool async) const {
    if (sl.inFlight) check(igab_patch_wait(h_));
    sl.block = block;
    sl.corners.clear();
    sl.inFlight = false;
    const int64_t e0 = block * blockElems_, e1 = std::min(nElements_, e0 + blockElems_);
    const igab_eval_out o = outOf(sl);
    if (async) {
      check(igab_eval_tensor_begin(h_, e0, e1, order_, nq1_.data(), xp_.data(), &o, nullptr));
      sl.inFlight = true;
    } else {
      check(igab_eval_tensor(h_, e0, e1, order_, nq1_.data(), xp_.data(), &o, IGAB_HOST, nullptr));
    }
    ++blockEvaluations_;
  }
  // the slot that holds element e, resident and complete; prefetches the following block
  Slot& slotOf(int64_t e) const {
    if (!haveRule_) throw Error(IGAB_INVALID_ARGUMENT, "setQuadratureRule first");
    const int64_t block = e / blockElems_;
    if (last_ && last_->block == block && !last_->inFlight) return *last_;
    Slot* sl = findSlot(block);
    if (!sl) {
      sl = &victim(-1);
      launch(*sl, block, false);
    } else if (sl->inFlight) {
      check(igab_patch_wait(h_));
      sl->inFlight = false;
    }
    sl->stamp = ++clock_;
    if (block + 1 < numBlocks_ && !findSlot(block + 1) && slots_.size() >= 2) {
      bool busy = false;
      for (auto& o : slots_) busy = busy || o.inFlight;
      if (!busy) {
        Slot& nx = victim(block);
        launch(nx, block + 1, true);
        nx.stamp = clock_;  // as recent as the block in use: the next miss evicts an older one
      }
    }
    last_ = sl;
    return *sl;
  }
  const double* field(int f, int64_t e, int64_t q, size_t width) const {
    const Slot& sl = slotOf(e);
    return sl.mem + fieldOff_[f] + ((size_t)(e - sl.block * blockElems_) * nq_ + q) * width;
  }
  void releaseWindow() {
    if (h_) igab_patch_wait(h_);
    for (auto& sl : slots_)
      if (sl.mem) igab_host_free(sl.mem);
    slots_.clear();
    last_ = nullptr;
    haveRule_ = false;
  }

  PatchData<dim, dimworld> data_;
  igab_patch* h_ = nullptr;
  int64_t nElements_ = 0, nDofs_ = 0, nq_ = 0;
  std::array<int32_t, 3> nElDir_{};
  int nb_ = 0, order_ = 1;
  bool haveRule_ = false;
  std::array<std::vector<double>, dim> xi_, w_;
  std::array<int, dim> nq1_{};
  std::array<const double*, dim> xp_{};
  std::vector<int32_t> dof_;
  // element-block window
  int64_t blockElems_ = 0, numBlocks_ = 0;
  size_t slotDoubles_ = 0, fieldOff_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  mutable std::vector<Slot> slots_;
  mutable Slot* last_ = nullptr;
  mutable uint64_t clock_ = 0;
  mutable int64_t blockEvaluations_ = 0;
};

// ---- NurbsLocalBasis -------------------------------------------------------------------------------------------------
template <int dim, int dimworld>
class LocalBasis {
public:
  LocalBasis() = default;
  LocalBasis(const Patch<dim, dimworld>* p, int64_t e) : p_(p), e_(e) {}

  // out[i] = R_i(in); Range is anything assignable from double through out[i][0] (FieldVector<R,1>) or out[i] (double)
  template <class Domain, class Range>
  void evaluateFunction(const Domain& in, std::vector<Range>& out) const {
    out.resize(p_->localSize());
    const int64_t q = p_->findRulePoint(in);
    if (q >= 0) {
      const double* r = p_->N(e_, q);
      for (int i = 0; i < p_->localSize(); ++i) set1(out[i], r[i]);
      return;
    }
    std::vector<double> tmp(p_->localSize());
    igab_eval_out o{};
    o.N = tmp.data();
    p_->evalPoint(e_, in, 0, o);
    for (int i = 0; i < p_->localSize(); ++i) set1(out[i], tmp[i]);
  }

  // out[i][0][j] = dR_i/dxi_j on the reference element (already scaled by the span length, nurbsbasis.hh:93-95)
  template <class Domain, class JacobianRange>
  void evaluateJacobian(const Domain& in, std::vector<JacobianRange>& out) const {
    out.resize(p_->localSize());
    const int64_t q = p_->findRulePoint(in);
    std::vector<double> tmp;
    const double* r;
    if (q >= 0) {
      r = p_->dN(e_, q);
    } else {
      tmp.resize((size_t)p_->localSize() * dim);
      igab_eval_out o{};
      o.dN = tmp.data();
      p_->evalPoint(e_, in, 1, o);
      r = tmp.data();
    }
    for (int i = 0; i < p_->localSize(); ++i)
      for (int j = 0; j < dim; ++j) out[i][0][j] = r[i * dim + j];
  }

  // any mixed derivative (nurbsbasis.hh:99-108); orders with |alpha| <= 2 on a rule point come from the cache
  template <class Domain, class Range>
  void partial(const std::array<unsigned int, dim>& order, const Domain& in, std::vector<Range>& out) const {
    out.resize(p_->localSize());
    unsigned total = 0;
    for (auto o : order) total += o;
    if (total == 0) return evaluateFunction(in, out);
    std::vector<double> tmp(p_->localSize());
    const int32_t el = (int32_t)e_;
    double xi[dim];
    int al[dim];
    for (int i = 0; i < dim; ++i) {
      xi[i] = in[i];
      al[i] = (int)order[i];
    }
    check(igab_partial(p_->handle(), 1, &el, xi, al, tmp.data(), IGAB_HOST, nullptr));
    for (int i = 0; i < p_->localSize(); ++i) set1(out[i], tmp[i]);
  }

  unsigned int order() const {
    int m = 0;
    for (int d : p_->patchData().degree) m = d > m ? d : m;
    return (unsigned)m;
  }
  std::size_t size() const { return (std::size_t)p_->localSize(); }

private:
  template <class T>
  static auto set1(T& t, double v) -> decltype(t[0] = v, void()) { t[0] = v; }
  static void set1(double& t, double v) { t = v; }
  const Patch<dim, dimworld>* p_ = nullptr;
  int64_t e_ = 0;
};

// ---- NURBSGeometry ---------------------------------------------------------------------------------------------------
template <int dim, int dimworld>
class Geometry {
public:
  using GlobalCoordinate = std::array<double, dimworld>;
  using JacobianTransposed = std::array<std::array<double, dimworld>, dim>;
  using JacobianInverseTransposed = std::array<std::array<double, dim>, dimworld>;
  Geometry() = default;
  Geometry(const Patch<dim, dimworld>* p, int64_t e) : p_(p), e_(e) {}

  template <class Local>
  GlobalCoordinate global(const Local& local) const {
    GlobalCoordinate r;
    const int64_t q = p_->findRulePoint(local);
    if (q >= 0) {
      for (int c = 0; c < dimworld; ++c) r[c] = p_->x(e_, q)[c];
    } else {
      igab_eval_out o{};
      o.x = r.data();
      p_->evalPoint(e_, local, 0, o);
    }
    return r;
  }
  template <class Local>
  JacobianTransposed jacobianTransposed(const Local& local) const {
    double buf[dim * dimworld];
    const double* s = buf;
    const int64_t q = p_->findRulePoint(local);
    if (q >= 0)
      s = p_->Jt(e_, q);
    else {
      igab_eval_out o{};
      o.Jt = buf;
      p_->evalPoint(e_, local, 1, o);
    }
    JacobianTransposed r;
    for (int d = 0; d < dim; ++d)
      for (int c = 0; c < dimworld; ++c) r[d][c] = s[d * dimworld + c];
    return r;
  }
  template <class Local>
  double integrationElement(const Local& local) const {
    const int64_t q = p_->findRulePoint(local);
    if (q >= 0) return p_->detJ(e_, q);
    double r = 0;
    igab_eval_out o{};
    o.detJ = &r;
    p_->evalPoint(e_, local, 1, o);
    return r;
  }
  template <class Local>
  JacobianInverseTransposed jacobianInverseTransposed(const Local& local) const {
    double buf[dim * dimworld];
    const double* s = buf;
    const int64_t q = p_->findRulePoint(local);
    if (q >= 0)
      s = p_->JinvT(e_, q);
    else {
      igab_eval_out o{};
      o.JinvT = buf;
      p_->evalPoint(e_, local, 1, o);
    }
    JacobianInverseTransposed r;
    for (int c = 0; c < dimworld; ++c)
      for (int d = 0; d < dim; ++d) r[c][d] = s[c * dim + d];
    return r;
  }
  // sum_q detJ w over the element's rule (nurbsgeometry.hh:96-113)
  double volume() const {
    double v = 0;
    for (int64_t q = 0; q < p_->pointsPerElement(); ++q) v += p_->detJ(e_, q) * p_->weight(q);
    return v;
  }
  // J = (J^T)^T and its pseudo-inverse (nurbsgeometry.hh:198,212-214)
  template <class Local>
  std::array<std::array<double, dim>, dimworld> jacobian(const Local& local) const {
    const auto jt = jacobianTransposed(local);
    std::array<std::array<double, dim>, dimworld> r;
    for (int c = 0; c < dimworld; ++c)
      for (int d = 0; d < dim; ++d) r[c][d] = jt[d][c];
    return r;
  }
  template <class Local>
  JacobianTransposed jacobianInverse(const Local& local) const {
    const auto jit = jacobianInverseTransposed(local);
    JacobianTransposed r;
    for (int d = 0; d < dim; ++d)
      for (int c = 0; c < dimworld; ++c) r[d][c] = jit[c][d];
    return r;
  }
  // rows [uu, vv(, ww), uv(, uw, vw)] of d^2 x / d xi^2 (nurbsgeometry.hh:257-292)
  using Hessian = std::array<std::array<double, dimworld>, dim*(dim + 1) / 2>;
  template <class Local>
  Hessian secondDerivativeOfPosition(const Local& local) const {
    constexpr int nh = dim * (dim + 1) / 2;
    double buf[nh * dimworld];
    const double* s = buf;
    const int64_t q = p_->derivOrder() >= 2 ? p_->findRulePoint(local) : -1;
    if (q >= 0)
      s = p_->H(e_, q);
    else {
      igab_eval_out o{};
      o.H = buf;
      p_->evalPoint(e_, local, 2, o);
    }
    Hessian r;
    for (int k = 0; k < nh; ++k)
      for (int c = 0; c < dimworld; ++c) r[k][c] = s[k * dimworld + c];
    return r;
  }
  // nurbsgeometry.hh:304-354
  template <class Local>
  std::tuple<GlobalCoordinate, JacobianTransposed, Hessian> zeroFirstAndSecondDerivativeOfPosition(
      const Local& local) const {
    return {global(local), jacobianTransposed(local), secondDerivativeOfPosition(local)};
  }

  // differential geometry of surfaces (nurbsgeometry.hh:216-255), evaluated on the device from J^T and H
  template <class Local>
  GlobalCoordinate normal(const Local& local) const {
    static_assert(dim == 2 && dimworld == 3, "normal: surfaces in R^3 only");
    GlobalCoordinate r;
    forms(local, false, r.data(), nullptr, nullptr, nullptr, nullptr);
    return r;
  }
  template <class Local>
  GlobalCoordinate unitNormal(const Local& local) const {
    static_assert(dim == 2 && dimworld == 3, "unitNormal: surfaces in R^3 only");
    GlobalCoordinate r;
    forms(local, false, nullptr, r.data(), nullptr, nullptr, nullptr);
    return r;
  }
  template <class Local>
  std::array<std::array<double, dim>, dim> metric(const Local& local) const {
    std::array<std::array<double, dim>, dim> r;
    forms(local, false, nullptr, nullptr, &r[0][0], nullptr, nullptr);
    return r;
  }
  template <class Local>
  std::array<std::array<double, 2>, 2> secondFundamentalForm(const Local& local) const {
    static_assert(dim == 2 && dimworld == 3, "secondFundamentalForm: surfaces in R^3 only");
    std::array<std::array<double, 2>, 2> r;
    forms(local, true, nullptr, nullptr, nullptr, &r[0][0], nullptr);
    return r;
  }
  template <class Local>
  double gaussianCurvature(const Local& local) const {
    static_assert(dim == 2 && dimworld == 3, "gaussianCurvature: surfaces in R^3 only");
    double r = 0;
    forms(local, true, nullptr, nullptr, nullptr, nullptr, &r);
    return r;
  }

  // inverse map (nurbsgeometry.hh:145-176): Newton for dim == dimworld, closest-point projection otherwise; the
  // reference's MathError ("no energy decreasing direction") becomes igab::Error
  template <class Global>
  std::array<double, dim> local(const Global& global) const {
    double X[dimworld];
    for (int c = 0; c < dimworld; ++c) X[c] = global[c];
    std::array<double, dim> r;
    const int32_t el = (int32_t)e_;
    int32_t flag = 0;
    check(igab_geometry_local(p_->handle(), 1, &el, X, r.data(), &flag, IGAB_HOST, nullptr));
    if (flag == 2) throw Error(IGAB_OK, "local: no energy-decreasing direction (Dune::MathError in the reference)");
    return r;
  }

  int corners() const { return 1 << dim; }
  GlobalCoordinate corner(int k) const {
    GlobalCoordinate r;
    if (const double* c = p_->cornerOrCenter(e_, k)) {
      for (int i = 0; i < dimworld; ++i) r[i] = c[i];
      return r;
    }
    std::array<double, dim> l;
    for (int i = 0; i < dim; ++i) l[i] = (k & (1 << i)) ? 1 : 0;
    return global(l);
  }
  GlobalCoordinate center() const {
    GlobalCoordinate r;
    if (const double* c = p_->cornerOrCenter(e_, 1 << dim)) {
      for (int i = 0; i < dimworld; ++i) r[i] = c[i];
      return r;
    }
    std::array<double, dim> l;
    l.fill(0.5);
    return global(l);
  }
  bool affine() const { return false; }

private:
  template <class Local>
  void forms(const Local& local, bool needH, double* n, double* un, double* g, double* b, double* K) const {
    const auto jt = jacobianTransposed(local);
    Hessian h{};
    if (needH) h = secondDerivativeOfPosition(local);
    check(igab_surface_forms(dim, dimworld, 1, &jt[0][0], needH ? &h[0][0] : nullptr, n, un, g, b, K, IGAB_HOST,
                             nullptr));
  }
  const Patch<dim, dimworld>* p_ = nullptr;
  int64_t e_ = 0;
};

// ---- NurbsLocalCoefficients (nurbsbasis.hh:150-290): LocalKey (subEntity, codim, index) per local shape function ----
struct LocalKey {
  unsigned int subEntity_, codim_, index_;
  unsigned int subEntity() const { return subEntity_; }
  unsigned int codim() const { return codim_; }
  unsigned int index() const { return index_; }
};
template <int dim>
class LocalCoefficients {
public:
  LocalCoefficients() = default;
  explicit LocalCoefficients(const std::array<int, dim>& degree) {
    std::size_t nb = 1;
    for (int d : degree) nb *= (std::size_t)d + 1;
    std::vector<uint32_t> s(nb), c(nb), x(nb);
    check(igab_local_keys(dim, degree.data(), s.data(), c.data(), x.data()));
    li_.resize(nb);
    for (std::size_t i = 0; i < nb; ++i) li_[i] = LocalKey{s[i], c[i], x[i]};
  }
  std::size_t size() const { return li_.size(); }
  const LocalKey& localKey(std::size_t i) const { return li_[i]; }

private:
  std::vector<LocalKey> li_;
};

// ---- LocalView of a NurbsBasis: bind(element) / index(i) / tree().finiteElement().localBasis() --------------------
template <int dim, int dimworld>
class LocalView {
public:
  explicit LocalView(std::shared_ptr<const Patch<dim, dimworld>> p)
      : p_(std::move(p)), lc_(p_->patchData().degree) {}
  void bind(int64_t elementIndex) {
    if (elementIndex < 0 || elementIndex >= p_->size0()) throw Error(IGAB_INVALID_ARGUMENT, "bind: no such element");
    e_ = elementIndex;
    bound_ = true;
    lb_ = LocalBasis<dim, dimworld>(p_.get(), e_);
  }
  void unbind() { bound_ = false; }
  std::size_t size() const { return (std::size_t)p_->localSize(); }
  // NurbsPreBasis::indices (nurbsbasis.hh:590-598): flat global index of local DOF i
  int32_t index(std::size_t i) const {
    if (!bound_) throw Error(IGAB_INVALID_ARGUMENT, "This element is not bound!");  // assert at nurbsbasis.hh:379
    return p_->dofsOf(e_)[i];
  }
  const LocalBasis<dim, dimworld>& localBasis() const {
    if (!bound_) throw Error(IGAB_INVALID_ARGUMENT, "This element is not bound!");
    return lb_;
  }
  // NurbsLocalFiniteElement::localCoefficients (nurbsbasis.hh:386-388)
  const LocalCoefficients<dim>& localCoefficients() const { return lc_; }
  Geometry<dim, dimworld> geometry() const { return Geometry<dim, dimworld>(p_.get(), e_); }

private:
  std::shared_ptr<const Patch<dim, dimworld>> p_;
  LocalCoefficients<dim> lc_;
  LocalBasis<dim, dimworld> lb_;
  int64_t e_ = 0;
  bool bound_ = false;
};

// ---- NurbsPreBasis (nurbsbasis.hh:503-700): sizes of the global basis and the local view factory --------------------
template <int dim, int dimworld>
class PreBasis {
public:
  explicit PreBasis(std::shared_ptr<const Patch<dim, dimworld>> p) : p_(std::move(p)) {}
  // number of basis functions of the (possibly trimmed) patch (nurbsbasis.hh:625-629)
  std::size_t size() const { return (std::size_t)p_->numDofs(); }
  std::size_t dimension() const { return size(); }
  // prod (p_d + 1) (nurbsbasis.hh:545-550)
  std::size_t maxNodeSize() const { return (std::size_t)p_->localSize(); }
  // n_d = size(U_d) - p_d - 1 (nurbsbasis.hh:631-633)
  unsigned int sizePerDirection(int d) const {
    const auto& pd = p_->patchData();
    return (unsigned int)(pd.knotSpans[d].size() - pd.degree[d] - 1);
  }
  LocalView<dim, dimworld> localView() const { return LocalView<dim, dimworld>(p_); }

private:
  std::shared_ptr<const Patch<dim, dimworld>> p_;
};

// ---- quadrature rules handed to the batched entry points (host-side; the C-ABI takes rules as input) -----------------
// Mirrors NURBSGridEntity<0>::fillQuadratureRule / getQuadratureRule (dune/iga/nurbsgridentity.hh:120-141) and
// fillQuadratureRuleImpl for trimmed elements (dune/iga/utils/fillquadraturerule.hh:8-19).  dune-geometry's tabulated
// Gauss rules are not vendored with the reference: Gauss-Legendre nodes are computed here (Newton on P_n).

// n-point Gauss-Legendre rule on [0,1], abscissae ascending
inline void gaussLegendre01(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign((size_t)n, 0.0);
  w.assign((size_t)n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 1.0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 1; j <= n; ++j) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j;
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[(size_t)i] = 0.5 * (1.0 - z);
    x[(size_t)(n - 1 - i)] = 0.5 * (1.0 + z);
    w[(size_t)i] = w[(size_t)(n - 1 - i)] = 1.0 / ((1.0 - z * z) * pp * pp);
  }
}

// Gauss points per direction for polynomial order `order` (dune-geometry: floor(order / 2) + 1)
inline int pointsForOrder(int order) { return order / 2 + 1; }

// the rule of an UNTRIMMED element: QuadratureRules<double,dim>::rule(cube, order), order defaulting to dim * max(p)
// (nurbsgridentity.hh:123-124,131-132)
template <int dim>
void getQuadratureRule(const std::array<int, dim>& degree, std::array<std::vector<double>, dim>& xi,
                       std::array<std::vector<double>, dim>& w, int order = -1) {
  int pmax = 0;
  for (int d = 0; d < dim; ++d) pmax = degree[d] > pmax ? degree[d] : pmax;
  const int n = pointsForOrder(order >= 0 ? order : dim * pmax);
  for (int d = 0; d < dim; ++d) gaussLegendre01(n, xi[d], w[d]);
}

// the rule of a TRIMMED element from its sub-triangles in element-local coordinates (fillquadraturerule.hh:13-19):
// position = triangle.global(ip), weight = ip.weight * triangle.integrationElement; 6-point rule of degree 4.
// triangles: [ntri][3][2]; appends to xi [npts][2], w [npts].
inline void fillQuadratureRule(const std::vector<double>& triangles, std::vector<double>& xi, std::vector<double>& w) {
  static const double rp[6][2] = {{0.445948490915965, 0.445948490915965}, {0.445948490915965, 0.108103018168070},
                                  {0.108103018168070, 0.445948490915965}, {0.091576213509771, 0.091576213509771},
                                  {0.091576213509771, 0.816847572980459}, {0.816847572980459, 0.091576213509771}};
  static const double rw[6] = {0.223381589678011 * 0.5, 0.223381589678011 * 0.5, 0.223381589678011 * 0.5,
                               0.109951743655322 * 0.5, 0.109951743655322 * 0.5, 0.109951743655322 * 0.5};
  for (size_t t = 0; t + 6 <= triangles.size(); t += 6) {
    const double* T = triangles.data() + t;
    const double ax = T[2] - T[0], ay = T[3] - T[1], bx = T[4] - T[0], by = T[5] - T[1];
    const double ie = std::fabs(ax * by - ay * bx);
    for (int k = 0; k < 6; ++k) {
      xi.push_back(T[0] + rp[k][0] * ax + rp[k][1] * bx);
      xi.push_back(T[1] + rp[k][0] * ay + rp[k][1] * by);
      w.push_back(rw[k] * ie);
    }
  }
}

// Gauss-Lobatto branch of fillQuadratureRuleImpl (fillquadraturerule.hh:21-44): points with exactly equal coordinates are
// merged (weights added in order of appearance) and the rule is returned sorted by (x, y).  In place.
inline void mergeDuplicatePoints(std::vector<double>& xi, std::vector<double>& w) {
  struct Less {
    bool operator()(const std::array<double, 2>& a, const std::array<double, 2>& b) const {
      if (a[0] < b[0]) return true;
      if (a[0] > b[0]) return false;
      return a[1] < b[1];
    }
  };
  std::map<std::array<double, 2>, double, Less> uniq;
  for (size_t k = 0; k < w.size(); ++k) {
    auto [it, fresh] = uniq.try_emplace({xi[2 * k], xi[2 * k + 1]}, w[k]);
    if (!fresh) it->second += w[k];
  }
  xi.clear();
  w.clear();
  for (const auto& [pos, weight] : uniq) {
    xi.push_back(pos[0]);
    xi.push_back(pos[1]);
    w.push_back(weight);
  }
}

// ---- multi-GPU: one process per GPU, element slabs, NCCL behind the C-ABI ------------------------------------------------
// The reference is sequential; a DUNE program that runs one rank per GPU creates the communicator once (the 128-byte id
// travels by MPI_Bcast or any other means the host program has) and hands it to the reductions / the assembler.
class Communicator {
public:
  using Id = std::array<char, 128>;
  static Id uniqueId() {
    Id id{};
    check(igab_comm_unique_id(id.data()));
    return id;
  }
  Communicator(int rank, int world, const Id& id) { check(igab_comm_create(rank, world, id.data(), &comm_)); }
  Communicator(const Communicator&) = delete;
  Communicator& operator=(const Communicator&) = delete;
  ~Communicator() { igab_comm_destroy(comm_); }
  void* get() const { return comm_; }
  int rank() const {
    int r = 0;
    check(igab_comm_info(comm_, &r, nullptr));
    return r;
  }
  int size() const {
    int w = 1;
    check(igab_comm_info(comm_, nullptr, &w));
    return w;
  }

private:
  void* comm_ = nullptr;
};

// [elem_begin, elem_end) of `rank`: whole layers of the slowest direction (igab_partition_slab)
template <int dim>
std::pair<int64_t, int64_t> partitionSlab(const std::array<int32_t, dim>& elementsPerDirection, int world, int rank) {
  int64_t b = 0, e = 0;
  check(igab_partition_slab(dim, elementsPerDirection.data(), world, rank, &b, &e));
  return {b, e};
}

// world + 1 element bounds of a trimmed patch balanced by the cumulative number of quadrature points (empty elements
// none, trimmed ones their own rules); perLayer > 1 rounds to whole layers of the slowest direction (igab_partition_weighted)
inline std::vector<int64_t> partitionWeighted(const std::vector<int64_t>& pointsPerElement, int world, int64_t perLayer = 1) {
  std::vector<int64_t> bounds((size_t)world + 1);
  check(igab_partition_weighted((int64_t)pointsPerElement.size(), pointsPerElement.data(), world, perLayer, bounds.data()));
  return bounds;
}

// ---- globalAssembler (dune/python/test/poisson.py:84-127), vectorised: the whole loop over elements, quadrature points
//      and local DOF pairs is one device pass; what comes back is the CSR matrix and the load vector -----------------------
struct CsrMatrix {
  int64_t rows = 0;
  std::vector<int64_t> rowptr;
  std::vector<int32_t> colidx;
  std::vector<double> values;
};

// own quadrature of the trimmed elements (NURBSGridEntity::getQuadratureRule, fillquadraturerule.hh:8-19):
// element elem[k] carries the points offset[k] .. offset[k+1] of xi [npts][2] (element-local) with weights w
struct TrimmedRules {
  std::vector<int32_t> elem;     // strictly ascending
  std::vector<int64_t> offset;   // elem.size() + 1 entries
  std::vector<double> xi, w;
};

template <int dim, int dimworld>
class GlobalAssembler {
public:
  // kind: IGAB_ASM_POISSON / IGAB_ASM_MASS (scalar basis) or IGAB_ASM_ELASTICITY (power basis, flat-interleaved)
  GlobalAssembler(std::shared_ptr<const Patch<dim, dimworld>> p, int kind)
      : p_(std::move(p)), kind_(kind), ncomp_(kind == IGAB_ASM_ELASTICITY ? dim : 1) {
    int64_t nnz = 0;
    check(igab_csr_pattern(p_->handle(), ncomp_, &A_.rows, &nnz, nullptr, nullptr, IGAB_HOST, nullptr));
    A_.rowptr.resize((size_t)A_.rows + 1);
    A_.colidx.resize((size_t)nnz);
    A_.values.assign((size_t)nnz, 0.0);
    check(igab_csr_pattern(p_->handle(), ncomp_, &A_.rows, &nnz, A_.rowptr.data(), A_.colidx.data(), IGAB_HOST, nullptr));
  }
  // A and b from the rule (xi, w per direction); D: the material matrix of elasticity (3x3 / 6x6, Voigt) or nullptr;
  // f: the constant source of poisson.py:79
  void assemble(const std::array<std::vector<double>, dim>& xi, const std::array<std::vector<double>, dim>& w,
                const double* D = nullptr, double f = 1.0, const TrimmedRules* trimmed = nullptr) {
    std::array<int, dim> nq1;
    std::array<const double*, dim> xp, wp;
    for (int i = 0; i < dim; ++i) {
      nq1[i] = (int)xi[i].size();
      xp[i] = xi[i].data();
      wp[i] = w[i].data();
    }
    b_.assign((size_t)A_.rows, 0.0);
    const int64_t nl = trimmed ? (int64_t)trimmed->elem.size() : 0;
    check(igab_global_assemble(p_->handle(), kind_, D, f, 0, p_->size0(), nq1.data(), xp.data(), wp.data(), nl,
                               nl ? trimmed->elem.data() : nullptr, nl ? trimmed->offset.data() : nullptr,
                               nl ? trimmed->xi.data() : nullptr, nl ? trimmed->w.data() : nullptr, A_.values.data(),
                               b_.data(), IGAB_HOST, nullptr));
  }
  // the same for the elements [elemBegin, elemEnd) only: one rank's slab of the shared pattern (partitionSlab); rows
  // that elements of several ranks touch hold partial sums until exchangeInterfaceRows has run
  void assemble(const std::array<std::vector<double>, dim>& xi, const std::array<std::vector<double>, dim>& w,
                int64_t elemBegin, int64_t elemEnd, const double* D = nullptr, double f = 1.0) {
    std::array<int, dim> nq1;
    std::array<const double*, dim> xp, wp;
    for (int i = 0; i < dim; ++i) {
      nq1[i] = (int)xi[i].size();
      xp[i] = xi[i].data();
      wp[i] = w[i].data();
    }
    b_.assign((size_t)A_.rows, 0.0);
    check(igab_global_assemble(p_->handle(), kind_, D, f, elemBegin, elemEnd, nq1.data(), xp.data(), wp.data(), 0, nullptr,
                               nullptr, nullptr, nullptr, A_.values.data(), b_.data(), IGAB_HOST, nullptr));
  }
  void exchangeInterfaceRows(const Communicator& comm) {
    check(igab_exchange_interface_rows(p_->handle(), ncomp_, nullptr, A_.values.data(), b_.data(), IGAB_HOST, comm.get(),
                                       nullptr));
  }
  // load vector from a source evaluated by the caller at the quadrature points (poisson.py:79 with f(posGlobal)):
  // f [n_elements * nq][ncomp], points in the order of the rule; replaces rhs()
  void assembleLoadVector(const std::array<std::vector<double>, dim>& xi, const std::array<std::vector<double>, dim>& w,
                          const std::vector<double>& f) {
    std::array<int, dim> nq1;
    std::array<const double*, dim> xp, wp;
    for (int i = 0; i < dim; ++i) {
      nq1[i] = (int)xi[i].size();
      xp[i] = xi[i].data();
      wp[i] = w[i].data();
    }
    b_.assign((size_t)A_.rows, 0.0);
    check(igab_global_load_vector(p_->handle(), ncomp_, 0, p_->size0(), nq1.data(), xp.data(), wp.data(), f.data(), 0,
                                  nullptr, nullptr, nullptr, nullptr, nullptr, b_.data(), IGAB_HOST, nullptr));
  }
  // rows of the marked DOFs become identity rows, b = g there (poisson.py:91-93,117-119)
  void applyDirichlet(const std::vector<uint8_t>& mask, const std::vector<double>* g = nullptr) {
    check(igab_csr_dirichlet(p_->handle(), ncomp_, mask.data(), g ? g->data() : nullptr, A_.values.data(), b_.data(),
                             IGAB_HOST, nullptr));
  }
  const CsrMatrix& matrix() const { return A_; }
  const std::vector<double>& rhs() const { return b_; }

private:
  std::shared_ptr<const Patch<dim, dimworld>> p_;
  int kind_, ncomp_;
  CsrMatrix A_;
  std::vector<double> b_;
};

}  // namespace igab
