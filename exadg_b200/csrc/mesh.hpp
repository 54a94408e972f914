// Host-side mesh description consumed by the GPU "MatrixFree replacement".
//
// What the reference gets from deal.II (Triangulation + Mapping + DoFHandler behind
// dealii::MatrixFree, set up in I/poisson/spatial_discretization/operator.cpp:261-284) is reduced
// here to what the SIPG Laplace data path needs: per locally relevant cell the MappingQ(m) support
// points, per owned cell the six face neighbours, and boundary types.  DoFs are numbered cell by
// cell in active-cell order (DG, no constraints: I/solvers_and_preconditioners/multigrid/constraints.h:120-125),
// so DoF index = cell index * (k+1)^3 + lexicographic local index.
//
// The hypercube generator restates, for the benchmark/test grids only,
//   I/grid/periodic_box.h:35-88           subdivided_hyper_cube(n_sub,-1,1) + periodic pairs + refine_global(l)
//   I/grid/deformed_cube_manifold.h:47-60  sine deformation of every support point
//   I/grid/grid_utilities.h:188-207        p4est partition: contiguous equal-count chunks of the
//                                          (coarse cell, Morton) curve
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <vector>

#include "tables.hpp"

namespace exadg_b200
{
enum BoundaryType : uint8_t { BT_INTERIOR = 0, BT_DIRICHLET = 1, BT_NEUMANN = 2 };

struct PeerPlan
{
  int rank = -1;
  std::vector<int32_t> send_cells; // local owned indices, ascending global id
  int64_t recv_begin = 0, recv_count = 0; // segment of the ghost range filled by this peer
};

struct HostMesh
{
  int mapping_degree = 1;
  int64_t n_owned = 0, n_ghost = 0;
  int64_t n_global_cells = 0, global_offset = 0;
  std::vector<double> xmap;    // [(owned+ghost)][(m+1)^3][3]
  std::vector<int32_t> nb;     // [owned][6]  local index (ghosts >= n_owned) or -1
  std::vector<uint8_t> nbface; // [owned][6]
  std::vector<uint8_t> bt;     // [(owned+ghost)][6]
  bool cartesian_uniform = false; // all cells are identical axis-aligned boxes
  double h[3] = {0, 0, 0};
  int rank = 0, world = 1;
  std::vector<PeerPlan> peers;
  std::vector<int64_t> ghost_global; // global cell id of each ghost
  int singular = -1; // 1: no Dirichlet face anywhere in the domain (OperatorBaseData::operator_is_singular, operator_base.h:59-103), 0: some, -1: derive from the local cells

  // unique faces touching owned cells
  int64_t n_faces = 0;
  std::vector<int32_t> face_id;    // [owned][6]
  std::vector<uint8_t> face_info;  // [owned][6]: bits 0-2 neighbour face, bit 3 = this cell is the plus side, bits 4-5 boundary type
  std::vector<int32_t> face_cells; // [n_faces][2] minus, plus (-1 on the boundary)
  std::vector<uint8_t> face_nos;   // [n_faces][2] face numbers on minus, plus side
  std::vector<uint8_t> face_bt;    // [n_faces]

  void build_faces()
  {
    face_id.assign(n_owned * 6, -1);
    face_info.assign(n_owned * 6, 0);
    face_cells.clear(); face_nos.clear(); face_bt.clear();
    n_faces = 0;
    for (int64_t c = 0; c < n_owned; ++c)
      for (int f = 0; f < 6; ++f) {
        const int64_t p = nb[c * 6 + f];
        const int fp = nbface[c * 6 + f];
        const uint8_t b = bt[c * 6 + f];
        bool create = true, plus = false;
        if (p >= 0 && p < n_owned) {
          if (p < c || (p == c && fp < f)) { create = false; plus = true; }
        }
        if (create) {
          face_id[c * 6 + f] = (int32_t)n_faces++;
          face_cells.push_back((int32_t)c); face_cells.push_back((int32_t)p);
          face_nos.push_back((uint8_t)f); face_nos.push_back((uint8_t)fp);
          face_bt.push_back(b);
        } else {
          face_id[c * 6 + f] = face_id[p * 6 + fp];
          if (face_id[c * 6 + f] < 0) throw std::runtime_error("inconsistent neighbour table (face not mutual)");
        }
        face_info[c * 6 + f] = (uint8_t)((fp & 7) | (plus ? 8 : 0) | ((b & 3) << 4));
      }
  }

  // all faces interior (periodic box) and a uniform Cartesian grid: the fast kernel applies
  bool all_interior() const
  {
    for (size_t i = 0; i < bt.size(); ++i) if (bt[i] != BT_INTERIOR) return false;
    for (int64_t i = 0; i < n_owned * 6; ++i) if (nb[i] < 0) return false;
    return true;
  }

  // constants lie in the kernel of the operator: periodic / Neumann faces only (operator_is_singular of the reference's operator data)
  bool pure_neumann_or_periodic() const
  {
    if (singular >= 0) return singular == 1;
    for (size_t i = 0; i < bt.size(); ++i) if (bt[i] == BT_DIRICHLET) return false;
    return true;
  }

  // structured orientation: neighbour across face f is entered through face f^1
  bool standard_orientation() const
  {
    for (int64_t c = 0; c < n_owned; ++c) for (int f = 0; f < 6; ++f)
      if (nb[c * 6 + f] >= 0 && nbface[c * 6 + f] != (f ^ 1)) return false;
    return true;
  }
};

struct HypercubeDesc
{
  int n_sub = 1, refine = 0, mapping_degree = 1;
  double deformation = 0.0; // 0: Cartesian
  int frequency = 2;
  int bc[6] = {0, 0, 0, 0, 0, 0}; // per domain face: 0 periodic, 1 Dirichlet, 2 Neumann
  int rank = 0, world = 1;
  double left = -1.0, right = 1.0;
};

struct HypercubeIndexer
{
  int n_sub, refine, n;
  HypercubeIndexer(int n_sub_, int refine_) : n_sub(n_sub_), refine(refine_), n(n_sub_ << refine_) {}
  int64_t n_cells() const { return (int64_t)n * n * n; }
  // active-cell order: coarse cells lexicographic (x fastest), descendants in z-order (child = x + 2y + 4z)
  void to_ijk(int64_t c, int ijk[3]) const
  {
    const int64_t per = (int64_t)1 << (3 * refine);
    const int64_t coarse = c / per; int64_t m = c % per;
    int x = 0, y = 0, z = 0;
    for (int l = 0; l < refine; ++l) {
      x |= (int)((m >> (3 * l)) & 1) << l; y |= (int)((m >> (3 * l + 1)) & 1) << l; z |= (int)((m >> (3 * l + 2)) & 1) << l;
    }
    ijk[0] = ((int)(coarse % n_sub) << refine) + x;
    ijk[1] = ((int)((coarse / n_sub) % n_sub) << refine) + y;
    ijk[2] = ((int)(coarse / ((int64_t)n_sub * n_sub)) << refine) + z;
  }
  int64_t to_cell(const int ijk[3]) const
  {
    const int mask = (1 << refine) - 1;
    int64_t m = 0;
    for (int l = 0; l < refine; ++l)
      m |= ((int64_t)((ijk[0] >> l) & 1) << (3 * l)) | ((int64_t)((ijk[1] >> l) & 1) << (3 * l + 1)) | ((int64_t)((ijk[2] >> l) & 1) << (3 * l + 2));
    (void)mask;
    const int64_t coarse = (ijk[0] >> refine) + (int64_t)n_sub * ((ijk[1] >> refine) + (int64_t)n_sub * (ijk[2] >> refine));
    return coarse * ((int64_t)1 << (3 * refine)) + m;
  }
};

inline HostMesh make_hypercube(const HypercubeDesc & d)
{
  if (d.n_sub < 1 || d.refine < 0 || d.mapping_degree < 1 || d.mapping_degree > 8) throw std::invalid_argument("bad hypercube parameters");
  if (d.world < 1 || d.rank < 0 || d.rank >= d.world) throw std::invalid_argument("bad rank/world");
  HypercubeIndexer ix(d.n_sub, d.refine);
  const int n = ix.n;
  const int64_t N = ix.n_cells();
  if (N * 1 > (int64_t)2000000000) throw std::invalid_argument("too many cells for 32-bit local indices");
  HostMesh M;
  M.mapping_degree = d.mapping_degree; M.rank = d.rank; M.world = d.world;
  M.n_global_cells = N;
  M.singular = 1;
  for (int f = 0; f < 6; ++f) if (d.bc[f] == BT_DIRICHLET) M.singular = 0;
  auto first_of = [&](int r) { return (int64_t)((__int128)N * r / d.world); };
  auto owner_of = [&](int64_t g) {
    int r = (int)(((__int128)g * d.world) / N);
    while (r + 1 < d.world && first_of(r + 1) <= g) ++r;
    while (r > 0 && first_of(r) > g) --r;
    return r;
  };
  const int64_t g0 = first_of(d.rank), g1 = first_of(d.rank + 1);
  M.global_offset = g0; M.n_owned = g1 - g0;

  // pass 1: neighbours (global ids), collect ghosts
  std::vector<int64_t> nbg(M.n_owned * 6);
  std::vector<uint8_t> bt_owned(M.n_owned * 6);
  std::vector<int64_t> ghosts;
  auto neighbour = [&](const int ijk[3], int f, int64_t & g, uint8_t & b) {
    const int dir = f / 2, s = f % 2;
    int q[3] = {ijk[0], ijk[1], ijk[2]};
    q[dir] += s ? 1 : -1;
    if (q[dir] < 0 || q[dir] >= n) {
      if (d.bc[f] == 0) { q[dir] = (q[dir] + n) % n; g = ix.to_cell(q); b = BT_INTERIOR; } // periodic = interior
      else { g = -1; b = (uint8_t)d.bc[f]; }
    } else { g = ix.to_cell(q); b = BT_INTERIOR; }
  };
  for (int64_t c = 0; c < M.n_owned; ++c) {
    int ijk[3]; ix.to_ijk(g0 + c, ijk);
    for (int f = 0; f < 6; ++f) {
      int64_t g; uint8_t b; neighbour(ijk, f, g, b);
      nbg[c * 6 + f] = g; bt_owned[c * 6 + f] = b;
      if (g >= 0 && (g < g0 || g >= g1)) ghosts.push_back(g);
    }
  }
  std::sort(ghosts.begin(), ghosts.end());
  ghosts.erase(std::unique(ghosts.begin(), ghosts.end()), ghosts.end());
  M.n_ghost = (int64_t)ghosts.size();
  M.ghost_global = ghosts;

  const int64_t nloc = M.n_owned + M.n_ghost;
  M.nb.resize(M.n_owned * 6); M.nbface.resize(M.n_owned * 6); M.bt.resize(nloc * 6);
  for (int64_t c = 0; c < M.n_owned; ++c)
    for (int f = 0; f < 6; ++f) {
      const int64_t g = nbg[c * 6 + f];
      int64_t l = -1;
      if (g >= g0 && g < g1) l = g - g0;
      else if (g >= 0) l = M.n_owned + (std::lower_bound(ghosts.begin(), ghosts.end(), g) - ghosts.begin());
      M.nb[c * 6 + f] = (int32_t)l;
      M.nbface[c * 6 + f] = (uint8_t)(f ^ 1);
      M.bt[c * 6 + f] = bt_owned[c * 6 + f];
    }

  // mapping support points and boundary types for all locally relevant cells
  const int np = d.mapping_degree + 1, np3 = np * np * np;
  std::vector<real_t> gl; lobatto_points(np, gl);
  const double hh = (d.right - d.left) / n;
  M.xmap.resize((size_t)nloc * np3 * 3);
  const double pi = 3.14159265358979323846;
  for (int64_t c = 0; c < nloc; ++c) {
    const int64_t g = c < M.n_owned ? g0 + c : ghosts[c - M.n_owned];
    int ijk[3]; ix.to_ijk(g, ijk);
    if (c >= M.n_owned)
      for (int f = 0; f < 6; ++f) { int64_t gg; uint8_t b; neighbour(ijk, f, gg, b); M.bt[c * 6 + f] = b; }
    for (int a2 = 0; a2 < np; ++a2) for (int a1 = 0; a1 < np; ++a1) for (int a0 = 0; a0 < np; ++a0) {
      const int a[3] = {a0, a1, a2};
      double X[3];
      for (int e = 0; e < 3; ++e)
        X[e] = (a[e] == np - 1) ? d.left + hh * (ijk[e] + 1) : d.left + hh * (ijk[e] + (double)gl[a[e]]);
      double * x = &M.xmap[((size_t)c * np3 + a0 + np * (a1 + np * a2)) * 3];
      if (d.deformation != 0.0) {
        double sinval = d.deformation; // deformed_cube_manifold.h:50-57
        for (int e = 0; e < 3; ++e) sinval *= std::sin(d.frequency * pi * (X[e] - d.left) / (d.right - d.left));
        for (int e = 0; e < 3; ++e) x[e] = X[e] + sinval;
      } else { x[0] = X[0]; x[1] = X[1]; x[2] = X[2]; }
    }
  }
  M.cartesian_uniform = (d.deformation == 0.0);
  M.h[0] = M.h[1] = M.h[2] = hh;

  // exchange plan: owned cells with a neighbour on rank p are sent to p; ghosts are grouped by owner
  if (d.world > 1) {
    std::map<int, std::vector<int32_t>> sends;
    for (int64_t c = 0; c < M.n_owned; ++c) {
      int seen[6]; int ns = 0;
      for (int f = 0; f < 6; ++f) {
        const int64_t g = nbg[c * 6 + f];
        if (g < 0 || (g >= g0 && g < g1)) continue;
        const int r = owner_of(g);
        bool dup = false;
        for (int i = 0; i < ns; ++i) dup |= (seen[i] == r);
        if (!dup) { seen[ns++] = r; sends[r].push_back((int32_t)c); }
      }
    }
    std::map<int, std::pair<int64_t, int64_t>> recvs;
    for (int64_t i = 0; i < M.n_ghost; ++i) {
      const int r = owner_of(ghosts[i]);
      auto it = recvs.find(r);
      if (it == recvs.end()) recvs[r] = std::make_pair(i, (int64_t)1); else it->second.second++;
    }
    for (auto & kv : sends) {
      PeerPlan p; p.rank = kv.first; p.send_cells = kv.second;
      auto it = recvs.find(kv.first);
      if (it == recvs.end()) throw std::runtime_error("asymmetric halo plan");
      p.recv_begin = it->second.first; p.recv_count = it->second.second;
      M.peers.push_back(p);
    }
    if (recvs.size() != sends.size()) throw std::runtime_error("asymmetric halo plan");
  }
  M.build_faces();
  return M;
}

// true if every cell of a degree-1 mapping is the same axis-aligned box
inline bool detect_cartesian_uniform(const HostMesh & M, double h[3])
{
  if (M.mapping_degree != 1) return false;
  const int64_t nloc = M.n_owned + M.n_ghost;
  if (nloc == 0) return false;
  for (int e = 0; e < 3; ++e) h[e] = M.xmap[(size_t)(1 << e) * 3 + e] - M.xmap[e];
  for (int64_t c = 0; c < nloc; ++c) {
    const double * X = &M.xmap[(size_t)c * 8 * 3];
    for (int v = 0; v < 8; ++v) for (int e = 0; e < 3; ++e) {
      const double expect = X[e] + (((v >> e) & 1) ? h[e] : 0.0);
      if (std::fabs(X[v * 3 + e] - expect) > 1e-13 * (std::fabs(h[e]) + std::fabs(expect))) return false;
    }
  }
  return h[0] > 0 && h[1] > 0 && h[2] > 0;
}

} // namespace exadg_b200
