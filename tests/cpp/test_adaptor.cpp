// Reads like the reference's own quadrature-loop tests (dune/iga/test/gridtests.cpp:338-360, 514-576 and
// dune/python/test/poisson.py:37-81), but through the batched adaptor over the C-ABI.  Checks every value against the
// plain-C oracle (oracle/igab_oracle.c, test infrastructure) linked into THIS test binary only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../dune-iga_b200/host/igab200_adaptor.hh"

extern "C" {
typedef struct {
  int dim, dimworld;
  int degree[3];
  int nknots[3];
  const double* knots[3];
  int ncp[3];
  const double* cp;
} igo_patch;
int igo_eval_tensor(const igo_patch*, long, long, int, const int*, const double*, double*, double*, double*, double*,
                    double*, double*, double*, double*, int);
long igo_dofmap(const igo_patch*, const unsigned char*, int*);
void igo_gauss01(int, double*, double*);
void igo_partial(const igo_patch*, long, const int*, const double*, const int*, double*);
int igo_assemble(const igo_patch*, int, const double*, double, long, long, const int*, const double*, const double*,
                 double*, double*, int);
}

static int failures = 0;
#define CHECK(c, ...)                   \
  do {                                  \
    if (!(c)) {                         \
      ++failures;                       \
      std::printf("FAIL %s:%d ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);         \
      std::printf("\n");                \
    }                                   \
  } while (0)

static bool close(double a, double b, double scale) { return std::fabs(a - b) <= 1e-12 * std::fmax(scale, 1e-300); }

int main() {
  constexpr int dim = 2, dw = 2;
  // quarter plate with hole, p = (2,2), 2 x 1 elements (SURVEY.md 8d C1 coarse patch)
  const double s = (1.0 + 1.0 / std::sqrt(2.0)) / 2.0, r2 = std::sqrt(2.0);
  igab::PatchData<dim, dw> pd;
  pd.degree = {2, 2};
  pd.knotSpans = {std::vector<double>{0, 0, 0, .5, 1, 1, 1}, std::vector<double>{0, 0, 0, 1, 1, 1}};
  pd.controlPoints = {-1, 0, 1,  -1, r2 - 1, s,  1 - r2, 1, s,  0, 1, 1,   -2.5, 0, 1,  -2.5, .75, 1,  -.75, 2.5, 1,
                      0, 2.5, 1,   -4, 0, 1,  -4, 4, 1,  -4, 4, 1,  0, 4, 1};
  auto patch = std::make_shared<igab::Patch<dim, dw>>(pd);
  CHECK(patch->size0() == 2, "elements %ld", (long)patch->size0());
  CHECK(patch->numDofs() == 12, "dofs %ld", (long)patch->numDofs());

  std::array<std::vector<double>, dim> xi, w;
  for (int d = 0; d < dim; ++d) {
    xi[d].resize(3);
    w[d].resize(3);
    igo_gauss01(3, xi[d].data(), w[d].data());
  }
  {  // the adaptor's own rules: Gauss-Legendre on [0,1] against the oracle's, the default order dim * max(p), a triangle
    std::array<std::vector<double>, dim> gx, gw;
    igab::getQuadratureRule<dim>(pd.degree, gx, gw);  // order 4 -> 3 points per direction
    CHECK(gx[0].size() == 3 && gx[1].size() == 3, "default rule size %zu", gx[0].size());
    for (int k = 0; k < 3; ++k)
      CHECK(std::fabs(gx[0][k] - xi[0][k]) < 1e-15 && std::fabs(gw[1][k] - w[1][k]) < 1e-15, "Gauss point %d", k);
    std::vector<double> g7x, g7w, o7x(7), o7w(7);
    igab::gaussLegendre01(7, g7x, g7w);
    igo_gauss01(7, o7x.data(), o7w.data());
    for (int k = 0; k < 7; ++k) CHECK(std::fabs(g7x[k] - o7x[k]) < 1e-15 && std::fabs(g7w[k] - o7w[k]) < 1e-15, "7-point rule %d", k);
    std::vector<double> tx, tw;
    igab::fillQuadratureRule({0.0, 0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0, 1.0}, tx, tw);  // the unit square in 2 triangles
    double area = 0, mx = 0;
    for (std::size_t k = 0; k < tw.size(); ++k) { area += tw[k]; mx += tw[k] * tx[2 * k] * tx[2 * k] * tx[2 * k + 1]; }
    CHECK(tw.size() == 12 && std::fabs(area - 1.0) < 1e-14 && std::fabs(mx - 1.0 / 6.0) < 1e-13, "triangle rule: area %g, int x^2 y %g", area, mx);
  }
  patch->setQuadratureRule(xi, w, 1);

  // oracle
  igo_patch op{};
  op.dim = dim; op.dimworld = dw;
  for (int d = 0; d < dim; ++d) {
    op.degree[d] = pd.degree[d];
    op.nknots[d] = (int)pd.knotSpans[d].size();
    op.knots[d] = pd.knotSpans[d].data();
    op.ncp[d] = op.nknots[d] - op.degree[d] - 1;
  }
  op.cp = pd.controlPoints.data();
  const int nq1[2] = {3, 3};
  std::vector<double> xcat(xi[0]);
  xcat.insert(xcat.end(), xi[1].begin(), xi[1].end());
  std::vector<double> wcat(w[0]);
  wcat.insert(wcat.end(), w[1].begin(), w[1].end());
  const int nb = 9, nq = 9, ne = 2;
  std::vector<double> oN(ne * nq * nb), odN(ne * nq * nb * 2), ox(ne * nq * 2), oJ(ne * nq * 4), odet(ne * nq), oJi(ne * nq * 4);
  igo_eval_tensor(&op, 0, ne, 1, nq1, xcat.data(), oN.data(), odN.data(), nullptr, ox.data(), oJ.data(), nullptr,
                  odet.data(), oJi.data(), 1);
  std::vector<int> odof(ne * nb);
  igo_dofmap(&op, nullptr, odof.data());
  std::vector<double> oKe(ne * nb * nb), ofe(ne * nb);
  igo_assemble(&op, 0, nullptr, 1.0, 0, ne, nq1, xcat.data(), wcat.data(), oKe.data(), ofe.data(), 1);

  // the loop of poisson.py:37-127, written against the adaptor
  igab::LocalView<dim, dw> localView(patch);
  double area = 0;
  const std::size_t nd = (std::size_t)patch->numDofs();
  std::vector<double> Aglob(nd * nd, 0.0), bglob(nd, 0.0);  // the scatter of poisson.py:113-125, dense for the check
  for (int64_t e = 0; e < patch->size0(); ++e) {
    localView.bind(e);
    const auto& localBasis = localView.localBasis();
    const auto geo = localView.geometry();
    const std::size_t n = localBasis.size();
    std::vector<double> localA(n * n, 0.0), localB(n, 0.0);
    for (int64_t q = 0; q < patch->pointsPerElement(); ++q) {
      std::array<double, dim> pos;
      patch->position(q, pos);
      std::vector<std::array<double, 1>> phi;
      std::vector<std::array<std::array<double, dim>, 1>> phiRefGradient;
      localBasis.evaluateFunction(pos, phi);
      localBasis.evaluateJacobian(pos, phiRefGradient);
      const auto JinvT = geo.jacobianInverseTransposed(pos);
      const double integrationElement = geo.integrationElement(pos);
      const auto posGlobal = geo.global(pos);
      const auto Jt = geo.jacobianTransposed(pos);
      const long pt = e * nq + q;
      for (std::size_t i = 0; i < n; ++i) {
        CHECK(close(phi[i][0], oN[pt * nb + i], 1.0), "N e=%ld q=%ld i=%zu", (long)e, (long)q, i);
        for (int j = 0; j < dim; ++j)
          CHECK(close(phiRefGradient[i][0][j], odN[(pt * nb + i) * 2 + j], 4.0), "dN");
      }
      for (int c = 0; c < dw; ++c) CHECK(close(posGlobal[c], ox[pt * 2 + c], 4.0), "x");
      for (int d = 0; d < dim; ++d)
        for (int c = 0; c < dw; ++c) {
          CHECK(close(Jt[d][c], oJ[pt * 4 + d * 2 + c], 8.0), "Jt");
          CHECK(close(JinvT[c][d], oJi[pt * 4 + c * 2 + d], 8.0), "JinvT");
        }
      CHECK(close(integrationElement, odet[pt], 8.0), "detJ");
      std::vector<std::array<double, dim>> grad(n);
      for (std::size_t i = 0; i < n; ++i)
        for (int c = 0; c < dw; ++c) {
          grad[i][c] = 0;
          for (int d = 0; d < dim; ++d) grad[i][c] += JinvT[c][d] * phiRefGradient[i][0][d];
        }
      const double wq = patch->weight(q);
      for (std::size_t i = 0; i < n; ++i) {
        for (std::size_t j = 0; j < n; ++j)
          localA[i * n + j] += wq * integrationElement * (grad[i][0] * grad[j][0] + grad[i][1] * grad[j][1]);
        localB[i] += wq * integrationElement * phi[i][0] * 1.0;
      }
    }
    for (std::size_t i = 0; i < n; ++i) {
      CHECK(localView.index(i) == odof[e * nb + i], "dof map e=%ld i=%zu", (long)e, i);
      CHECK(close(localB[i], ofe[e * nb + i], 1.0), "fe");
      for (std::size_t j = 0; j < n; ++j) CHECK(close(localA[i * n + j], oKe[(e * nb + i) * nb + j], 10.0), "Ke");
    }
    for (std::size_t i = 0; i < n; ++i) {
      bglob[(std::size_t)localView.index(i)] += localB[i];
      for (std::size_t j = 0; j < n; ++j)
        Aglob[(std::size_t)localView.index(i) * nd + (std::size_t)localView.index(j)] += localA[i * n + j];
    }
    area += geo.volume();
  }
  {  // the same system from the vectorised globalAssembler: one device pass, CSR back
    igab::GlobalAssembler<dim, dw> assembler(patch, IGAB_ASM_POISSON);
    assembler.assemble(xi, w, nullptr, 1.0);
    const igab::CsrMatrix& A = assembler.matrix();
    CHECK(A.rows == (int64_t)nd, "global rows %ld", (long)A.rows);
    std::vector<double> dense(nd * nd, 0.0);
    for (int64_t r = 0; r < A.rows; ++r)
      for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k) dense[(std::size_t)r * nd + A.colidx[k]] += A.values[k];
    double amax = 0;
    for (double v : Aglob) amax = std::fmax(amax, std::fabs(v));
    for (std::size_t k = 0; k < nd * nd; ++k) CHECK(close(dense[k], Aglob[k], amax), "global A entry %zu: %g vs %g", k, dense[k], Aglob[k]);
    for (std::size_t k = 0; k < nd; ++k) CHECK(close(assembler.rhs()[k], bglob[k], 1.0), "global b entry %zu", k);
    // Dirichlet rows as poisson.py:117-119: identity row, b = g
    std::vector<uint8_t> mask(nd, 0);
    std::vector<double> g(nd, 7.0);
    mask[0] = mask[nd - 1] = 1;
    assembler.applyDirichlet(mask, &g);
    const igab::CsrMatrix& AD = assembler.matrix();
    for (int64_t r : {(int64_t)0, (int64_t)nd - 1}) {
      for (int64_t k = AD.rowptr[r]; k < AD.rowptr[r + 1]; ++k)
        CHECK(AD.values[k] == (AD.colidx[k] == r ? 1.0 : 0.0), "dirichlet row %ld", (long)r);
      CHECK(assembler.rhs()[(std::size_t)r] == 7.0, "dirichlet rhs");
    }
    CHECK(close(assembler.rhs()[1], bglob[1], 1.0), "untouched rhs");
  }
  CHECK(std::fabs(area - patch->volume()) < 1e-13, "volume sum %g vs reduction %g", area, patch->volume());
  CHECK(std::fabs(area - (16.0 - M_PI / 4.0)) < 2e-2, "area %g", area);  // coarse 2-element patch, 3x3 Gauss

  // off-rule points: corners and corner +- 1e-8 (gridtests.cpp:514-576): evaluateJacobian[:,d] == partial(e_d) to 1e-14
  localView.bind(1);
  const double probes[5] = {0.0, 1e-8, 0.3, 1 - 1e-8, 1.0};
  for (double a : probes)
    for (double b : probes) {
      std::array<double, dim> pos{a, b};
      std::vector<std::array<std::array<double, dim>, 1>> J;
      localView.localBasis().evaluateJacobian(pos, J);
      for (int d = 0; d < dim; ++d) {
        std::array<unsigned int, dim> order{};
        order[d] = 1;
        std::vector<std::array<double, 1>> part;
        localView.localBasis().partial(order, pos, part);
        for (std::size_t i = 0; i < J.size(); ++i)
          CHECK(std::fabs(J[i][0][d] - part[i][0]) <= 1e-14 * std::fmax(1.0, std::fabs(part[i][0])), "J vs partial");
      }
    }
  CHECK(localView.localBasis().order() == 2 && localView.localBasis().size() == 9, "order/size");
  {
    igab::PreBasis<dim, dw> pre(patch);
    CHECK(pre.size() == (std::size_t)patch->numDofs() && pre.dimension() == pre.size() && pre.maxNodeSize() == 9, "prebasis sizes");
    CHECK(pre.sizePerDirection(0) * pre.sizePerDirection(1) == pre.size(), "sizePerDirection");
    auto lv = pre.localView();
    lv.bind(0);
    CHECK(lv.size() == 9 && lv.index(3) == patch->dofsOf(0)[3], "localView from prebasis");
  }
  {  // LocalKeys of the p = (2,2) element (nurbsbasis.hh:182-216): corners codim 2, edge midpoints codim 1, centre codim 0
    const auto& lc = localView.localCoefficients();
    const unsigned sub[9] = {0, 2, 1, 0, 0, 1, 2, 3, 3}, cod[9] = {2, 1, 2, 1, 0, 1, 2, 1, 2};
    CHECK(lc.size() == 9, "local coefficients size");
    for (int i = 0; i < 9; ++i)
      CHECK(lc.localKey(i).subEntity() == sub[i] && lc.localKey(i).codim() == cod[i] && lc.localKey(i).index() == 0, "local key %d", i);
  }

  // error behaviour: unbound view, bad element, bad knot vector
  igab::LocalView<dim, dw> v2(patch);
  bool threw = false;
  try { v2.index(0); } catch (const igab::Error&) { threw = true; }
  CHECK(threw, "unbound index must throw");
  threw = false;
  try { v2.bind(99); } catch (const igab::Error&) { threw = true; }
  CHECK(threw, "bind out of range must throw");
  threw = false;
  try {
    auto bad = pd;
    bad.knotSpans[0] = {0, 0, .5, 1, 1, 1, 1};
    igab::Patch<dim, dw> pbad(bad);
  } catch (const igab::Error& e) { threw = e.status == IGAB_INVALID_ARGUMENT; }
  CHECK(threw, "closed knot vector must be rejected");

  // Geometry::local round trip on the plate (gridtests.cpp:1085-1100, away from the degenerate corner)
  {
    igab::Geometry<dim, dw> geo(patch.get(), 0);
    const std::array<double, dim> xi{0.3, 0.6};
    const auto X = geo.global(xi);
    const auto back = geo.local(X);
    CHECK(std::fabs(back[0] - xi[0]) < 1e-13 && std::fabs(back[1] - xi[1]) < 1e-13, "local(global(xi)) %g %g", back[0], back[1]);
    const auto J = geo.jacobian(xi);
    const auto Jt2 = geo.jacobianTransposed(xi);
    CHECK(J[1][0] == Jt2[0][1], "jacobian is the transpose");
    const auto g = geo.metric(xi);
    CHECK(std::fabs(g[0][1] - (Jt2[0][0] * Jt2[1][0] + Jt2[0][1] * Jt2[1][1])) < 1e-13, "metric");
  }
  // surface of trimmedgridtests.cpp:112-186: literals of local(), plus the surface forms
  {
    igab::PatchData<2, 3> sd;
    sd.degree = {2, 2};
    sd.knotSpans[0] = sd.knotSpans[1] = {0, 0, 0, 1, 1, 1};
    const double tab[3][3][3] = {{{0, 0, 1}, {1, 0, 1}, {2, 0, 2}}, {{0, 1, 0}, {1, 1, 0}, {2, 1, 0}}, {{0, 2, 1}, {1, 2, 2}, {2, 2, 2}}};
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) {
        for (int c = 0; c < 3; ++c) sd.controlPoints.push_back(tab[i][j][c]);
        sd.controlPoints.push_back(1.0);
      }
    auto sp = std::make_shared<igab::Patch<2, 3>>(sd);
    igab::Geometry<2, 3> geo(sp.get(), 0);
    const auto u1 = geo.local(std::array<double, 3>{1.0, 1.0, 0.75});
    CHECK(std::fabs(u1[0] - 0.5) < 1e-14 && std::fabs(u1[1] - 0.5) < 1e-14, "surface local u1 %g %g", u1[0], u1[1]);
    const auto u2 = geo.local(std::array<double, 3>{2, 0, 2});
    CHECK(u2[0] == 0.0 && u2[1] == 1.0, "surface local u2 %g %g", u2[0], u2[1]);
    const std::array<double, 2> mid{0.5, 0.5};
    const auto Jt = geo.jacobianTransposed(mid);  // {0,2,.5},{2,0,.5}
    const auto N = geo.normal(mid);
    CHECK(std::fabs(N[0] - 1.0) < 1e-14 && std::fabs(N[1] - 1.0) < 1e-14 && std::fabs(N[2] + 4.0) < 1e-14, "normal %g %g %g", N[0], N[1], N[2]);
    const auto n = geo.unitNormal(mid);
    CHECK(std::fabs(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] - 1.0) < 1e-15, "unit normal");
    CHECK(std::fabs(n[0] * Jt[0][0] + n[1] * Jt[0][1] + n[2] * Jt[0][2]) < 1e-15, "normal orthogonal to tangent");
    const auto H = geo.secondDerivativeOfPosition(mid);
    const auto b = geo.secondFundamentalForm(mid);
    CHECK(std::fabs(b[0][1] - (H[2][0] * n[0] + H[2][1] * n[1] + H[2][2] * n[2])) < 1e-14 && b[0][1] == b[1][0], "second fundamental form");
    const auto g = geo.metric(mid);
    const double K = geo.gaussianCurvature(mid);
    CHECK(std::fabs(K - (b[0][0] * b[1][1] - b[0][1] * b[1][0]) / (g[0][0] * g[1][1] - g[0][1] * g[1][0])) < 1e-14, "gaussian curvature");
    auto [x0, J0, H0] = geo.zeroFirstAndSecondDerivativeOfPosition(mid);
    CHECK(std::fabs(x0[2] - 0.75) < 1e-15 && J0[0][1] == Jt[0][1] && H0[0][0] == H[0][0], "zeroFirstAndSecondDerivativeOfPosition");
  }

  // ---- real <-> direct element indices on a trimmed patch: the pattern of trimmedgridtests.cpp:335-373 (element 0 empty,
  //      1..3 trimmed: real(i) = i - 1, direct(i - 1) = i, getRealIndex(0) throws), the Gauss-Lobatto merge of trimmed
  //      rules (fillquadraturerule.hh:21-44), a load vector from a per-point source, a single-rank communicator
  {
    igab::PatchData<2, 2> td;
    td.degree = {1, 1};
    td.knotSpans = {std::vector<double>{0, 0, .5, 1, 1}, std::vector<double>{0, 0, .5, 1, 1}};
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) {
        td.controlPoints.push_back(0.5 * i);
        td.controlPoints.push_back(0.5 * j);
        td.controlPoints.push_back(1.0);
      }
    const std::vector<uint8_t> flags = {IGAB_TRIM_EMPTY, IGAB_TRIM_TRIMMED, IGAB_TRIM_TRIMMED, IGAB_TRIM_TRIMMED};
    igab::Patch<2, 2> tp(td, &flags);
    for (int i = 1; i < 4; ++i) CHECK(tp.getRealIndex(i) == i - 1 && tp.getDirectIndex(i - 1) == i, "real / direct index %d", i);
    bool threw = false;
    try { tp.getRealIndex(0); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw && tp.numRealElements() == 3, "empty element has no real index");
    std::vector<double> mx = {0.5, 0.5, 0.0, 0.0, 0.5, 0.5, 0.25, 1.0, 0.0, 0.0}, mw = {1, 2, 3, 4, 5};
    igab::mergeDuplicatePoints(mx, mw);
    CHECK(mw.size() == 3 && mx[0] == 0.0 && mw[0] == 7.0 && mx[2] == 0.25 && mw[1] == 4.0 && mw[2] == 4.0, "Lobatto merge");
    // per-point source f = 1 + x on the plate patch: sum of the load vector = integral of f
    igab::GlobalAssembler<dim, dw> gf(patch, IGAB_ASM_POISSON);
    std::vector<double> fsrc((size_t)patch->size0() * 9);
    double integral = 0;
    for (int64_t e = 0; e < patch->size0(); ++e)
      for (int64_t q = 0; q < 9; ++q) {
        fsrc[(size_t)e * 9 + q] = 1.0 + patch->x(e, q)[0];
        integral += fsrc[(size_t)e * 9 + q] * patch->detJ(e, q) * patch->weight(q);
      }
    gf.assembleLoadVector(xi, w, fsrc);
    double bs = 0;
    for (double v : gf.rhs()) bs += v;
    CHECK(std::fabs(bs - integral) <= 1e-13 * std::fabs(integral), "load vector with a per-point source: %g vs %g", bs, integral);
    // a one-rank communicator: slab = everything, the exchange finds no peer and leaves the system alone
    igab::Communicator comm(0, 1, igab::Communicator::uniqueId());
    CHECK(comm.rank() == 0 && comm.size() == 1, "communicator");
    const auto slab = igab::partitionSlab<dim>({2, 1}, 1, 0);
    CHECK(slab.first == 0 && slab.second == 2, "slab");
    igab::GlobalAssembler<dim, dw> gs(patch, IGAB_ASM_POISSON);
    gs.assemble(xi, w, slab.first, slab.second);
    const std::vector<double> before = gs.matrix().values;
    gs.exchangeInterfaceRows(comm);
    CHECK(before == gs.matrix().values, "exchange without peers");
  }

  // ---- a patch whose results do not fit the window: the element loop of gridtests.cpp:338-341 over 48^3 (p = 1) and
  //      32^3 (p = 2, checked against the oracle on a sample) elements with a small window of element blocks in pinned
  //      host memory.  Host memory stays at the window size, every block is evaluated exactly once when the elements are
  //      walked in index order, the next block is in flight while the current one is consumed.
  {
    const char* big = std::getenv("IGAB_TEST_FULL");  // IGAB_TEST_FULL=1: the same walk over 128^3 elements (minutes)
    for (int pdeg = 1; pdeg <= 2; ++pdeg) {
      const int n = big ? 128 : (pdeg == 1 ? 48 : 32);
      igab::PatchData<3, 3> bd;
      bd.degree = {pdeg, pdeg, pdeg};
      for (int d = 0; d < 3; ++d) {
        for (int k = 0; k <= pdeg; ++k) bd.knotSpans[d].push_back(0.0);
        for (int k = 1; k < n; ++k) bd.knotSpans[d].push_back((double)k / n);
        for (int k = 0; k <= pdeg; ++k) bd.knotSpans[d].push_back(1.0);
      }
      const int ncp = n + pdeg;
      // a smoothly distorted brick [0,2] x [0,1] x [0,3] with non-uniform weights (Greville abscissae as parameters)
      auto grev = [&](int i) {
        double sum = 0;
        for (int k = 1; k <= pdeg; ++k) sum += bd.knotSpans[0][(size_t)(i + k)];
        return sum / pdeg;
      };
      for (int k = 0; k < ncp; ++k)
        for (int j = 0; j < ncp; ++j)
          for (int i = 0; i < ncp; ++i) {
            const double u = grev(i), v = grev(j), t = grev(k);
            bd.controlPoints.push_back(2.0 * u + 0.05 * std::sin(3.0 * v));
            bd.controlPoints.push_back(v + 0.04 * u * t);
            bd.controlPoints.push_back(3.0 * t + 0.03 * std::cos(2.0 * u));
            bd.controlPoints.push_back(1.0 + 0.2 * u * v);
          }
      auto bp = std::make_shared<igab::Patch<3, 3>>(bd);
      std::array<std::vector<double>, 3> bx, bw;
      igab::getQuadratureRule<3>(bd.degree, bx, bw, 2 * pdeg);  // p + 1 points per direction
      const size_t window = size_t(48) << 20;
      bp->setQuadratureRule(bx, bw, 1, window, 3);
      const int64_t nq = bp->pointsPerElement();
      CHECK(bp->windowBytes() <= window, "window %zu bytes", bp->windowBytes());
      const double full = (double)bp->size0() * nq * (bp->localSize() * 4 + 3 + 9 + 1 + 9) * 8.0;
      CHECK(full > 4.0 * window, "the test patch must not fit the window (%g bytes)", full);
      double vol = 0, pu = 0, worst = 0;
      igab::LocalView<3, 3> lv(bp);
      for (int64_t e = 0; e < bp->size0(); ++e) {
        lv.bind(e);
        const auto geo = lv.geometry();
        vol += geo.volume();
        if (e % 97 == 0) {  // partition of unity and sum of gradients at a rule point of a sample of elements
          std::array<double, 3> q;
          bp->position(nq / 2, q);
          std::vector<double> ph;
          std::vector<std::array<std::array<double, 3>, 1>> gr;
          lv.localBasis().evaluateFunction(q, ph);
          lv.localBasis().evaluateJacobian(q, gr);
          double s0 = 0, s1 = 0;
          for (size_t i = 0; i < ph.size(); ++i) { s0 += ph[i]; s1 += gr[i][0][0] + gr[i][0][1] + gr[i][0][2]; }
          pu = std::fmax(pu, std::fabs(s0 - 1.0));
          worst = std::fmax(worst, std::fabs(s1));
        }
      }
      const int64_t blocks = (bp->size0() + bp->blockElements() - 1) / bp->blockElements();
      CHECK(bp->blockEvaluations() == blocks, "blocks evaluated %ld of %ld", (long)bp->blockEvaluations(), (long)blocks);
      CHECK(pu < 1e-13 && worst < 1e-10, "partition of unity %g, gradient sum %g", pu, worst);
      const double dev = bp->volume();  // the device reduction over the same rule
      CHECK(std::fabs(vol - dev) <= 1e-12 * dev, "windowed element loop volume %.15g, device reduction %.15g", vol, dev);
      // random access far apart evicts and re-evaluates, values stay those of the oracle
      igo_patch bo{};
      bo.dim = 3; bo.dimworld = 3;
      for (int d = 0; d < 3; ++d) {
        bo.degree[d] = pdeg;
        bo.nknots[d] = (int)bd.knotSpans[d].size();
        bo.knots[d] = bd.knotSpans[d].data();
        bo.ncp[d] = ncp;
      }
      bo.cp = bd.controlPoints.data();
      const int bnq1[3] = {pdeg + 1, pdeg + 1, pdeg + 1};
      std::vector<double> bxc;
      for (int d = 0; d < 3; ++d) bxc.insert(bxc.end(), bx[d].begin(), bx[d].end());
      const int nb3 = bp->localSize();
      std::vector<double> rN((size_t)nq * nb3), rdN((size_t)nq * nb3 * 3), rx((size_t)nq * 3), rJ((size_t)nq * 9), rd((size_t)nq),
          rJi((size_t)nq * 9);
      const int64_t probe[5] = {bp->size0() - 1, 0, bp->size0() / 2, 3, bp->size0() - 2};
      for (int64_t e : probe) {
        igo_eval_tensor(&bo, (long)e, (long)e + 1, 1, bnq1, bxc.data(), rN.data(), rdN.data(), nullptr, rx.data(), rJ.data(),
                        nullptr, rd.data(), rJi.data(), 1);
        for (int64_t q = 0; q < nq; q += 3) {
          for (int i = 0; i < nb3; ++i) CHECK(close(bp->N(e, q)[i], rN[(size_t)q * nb3 + i], 1.0), "window N e=%ld", (long)e);
          for (int i = 0; i < nb3 * 3; ++i) CHECK(close(bp->dN(e, q)[i], rdN[(size_t)q * nb3 * 3 + i], 8.0), "window dN e=%ld", (long)e);
          CHECK(close(bp->detJ(e, q), rd[(size_t)q], rd[(size_t)q]), "window detJ e=%ld", (long)e);
        }
        // corners and centre of the element: one batched call per block, equal to global() at those points
        igab::Geometry<3, 3> g3(bp.get(), e);
        const auto c5 = g3.corner(5), ce = g3.center();
        const auto d5 = g3.global(std::array<double, 3>{1.0, 0.0, 1.0}), de = g3.global(std::array<double, 3>{0.5, 0.5, 0.5});
        for (int c = 0; c < 3; ++c)
          CHECK(close(c5[c], d5[c], 3.0) && close(ce[c], de[c], 3.0), "corner / centre e=%ld: %.17g %.17g | %.17g %.17g", (long)e,
                c5[c], d5[c], ce[c], de[c]);
      }
      CHECK(bp->blockEvaluations() > blocks && bp->windowBytes() <= window, "eviction");
      std::printf("windowed walk p=%d: %ld elements, %ld blocks of %ld, window %.1f MB of %.1f MB, volume %.12f\n", pdeg,
                  (long)bp->size0(), (long)blocks, (long)bp->blockElements(), bp->windowBytes() / 1048576.0, full / 1048576.0, vol);
    }
  }

  std::printf(failures ? "adaptor test: %d FAILURES\n" : "adaptor test: ok\n", failures);
  return failures ? 1 : 0;
}
