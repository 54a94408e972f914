// Host-side plan of the pipelined host-buffer vmult (exadg_b200_vmult_host_pipelined): PCIe is full duplex, so the upload of src
// and the download of dst can overlap inside one call if the operator is applied chunk by chunk.  The vector is cut into
// contiguous chunks of cells; a chunk can be applied as soon as the chunks holding the face neighbours of its cells have been
// uploaded (the SIPG operator couples a cell with its six neighbours only: face_loop of I/operators/operator_base.cpp:1372-1397).
// This header decides the upload order and the moment every chunk becomes computable; it is pure host code (CPU-testable through
// exadg_b200_host_pipeline_plan), the stream / event choreography is in c_api.cu.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdint>
#include <vector>

namespace exadg_b200
{
struct HostPipelinePlan
{
  int n_chunks = 0;                   // 0: not applicable (ghost neighbours: partitioned operators keep the sequential path)
  int64_t cells_per_chunk = 0;
  std::vector<int32_t> upload_order;  // chunk ids in the order they are uploaded
  std::vector<int32_t> compute_order; // chunk ids in the order they become computable
  std::vector<int32_t> ready_chunk;   // per chunk: the chunk whose upload completes its dependencies
};

// nb: [n_owned][6] neighbour cell (or < 0 on the boundary); cells_per_chunk must be a multiple of the kernels' batch size
inline HostPipelinePlan build_host_pipeline(const int32_t * nb, int64_t n_owned, int64_t cells_per_chunk)
{
  HostPipelinePlan P;
  if (n_owned <= 0 || cells_per_chunk <= 0) return P;
  const int K = (int)((n_owned + cells_per_chunk - 1) / cells_per_chunk);
  // dependency sets: chunk a needs chunk b if a cell of a has a neighbour in b
  std::vector<std::vector<int32_t>> deps(K);
  for (int a = 0; a < K; ++a) deps[a].push_back(a);
  for (int64_t c = 0; c < n_owned; ++c) {
    const int a = (int)(c / cells_per_chunk);
    for (int f = 0; f < 6; ++f) {
      const int32_t p = nb[c * 6 + f];
      if (p < 0) continue;
      if (p >= n_owned) return P; // ghost cell
      const int b = (int)(p / cells_per_chunk);
      if (b != a && (deps[a].empty() || deps[a].back() != b)) deps[a].push_back(b);
    }
  }
  std::vector<std::vector<int32_t>> needed_by(K);
  for (int a = 0; a < K; ++a) {
    std::sort(deps[a].begin(), deps[a].end());
    deps[a].erase(std::unique(deps[a].begin(), deps[a].end()), deps[a].end());
    for (int32_t b : deps[a]) needed_by[b].push_back(a);
  }
  // greedy upload order: complete the chunk that misses the fewest uploads, so that chunks become computable early and steadily
  P.n_chunks = K; P.cells_per_chunk = cells_per_chunk;
  P.ready_chunk.assign(K, -1);
  std::vector<int> missing(K);
  std::vector<char> uploaded(K, 0);
  for (int a = 0; a < K; ++a) missing[a] = (int)deps[a].size();
  auto upload = [&](int b) {
    if (uploaded[b]) return;
    uploaded[b] = 1; P.upload_order.push_back(b);
    for (int32_t a : needed_by[b])
      if (--missing[a] == 0) { P.ready_chunk[a] = b; P.compute_order.push_back(a); }
  };
  while ((int)P.upload_order.size() < K) {
    int best = -1;
    for (int a = 0; a < K; ++a)
      if (missing[a] > 0 && (best < 0 || missing[a] < missing[best])) best = a;
    if (best < 0) break; // cannot happen: a chunk that is not uploaded misses at least itself
    for (int32_t b : deps[best]) upload(b);
  }
  return P;
}

// model of the plan: duration of one call in units of the time one direction takes alone (uploads back to back, downloads in
// compute order as soon as a chunk is computable, compute time neglected); 2.0 = no overlap
inline double host_pipeline_model(const HostPipelinePlan & P)
{
  if (P.n_chunks == 0) return 2.0;
  std::vector<int> pos(P.n_chunks);
  for (int i = 0; i < P.n_chunks; ++i) pos[P.upload_order[i]] = i + 1; // time at which the upload of the chunk is complete
  double t = 0;
  for (int32_t c : P.compute_order) t = std::max(t, (double)pos[P.ready_chunk[c]]) + 1.0;
  return t / P.n_chunks;
}
// ---- second plan (round 2): pieces in address order, readiness tracked per kernel unit, results stored by the kernels themselves ----
// The chunk plan above treats a chunk as computable only when ALL chunks touching it have arrived and downloads it as a whole; on the
// Morton curve of a periodic box a few far neighbours (the wrap-around layers, the faces between coarse cells) then hold back whole
// chunks (1.46 transfer times per call at 12288 cells per chunk).  Here the vector is uploaded piece by piece in address order and
// every UNIT of the kernel (a batch of the affine kernels, a cell of the general kernel) is applied by the launch behind the first
// upload that completes its own cells and their face neighbours (face_loop of I/operators/operator_base.cpp:1372-1397 couples a cell
// with its six neighbours only).  The launches write dst straight into the caller's pinned host buffer (device-mapped; every unit is
// stored exactly once, by one bulk store or coalesced stores), so there is no download granularity to wait for: on the 96^3 periodic
// box 1.11 transfer times per call with 72 pieces.
struct HostStreamPlan
{
  int n_steps = 0;                  // 0: not applicable
  int unit = 1;                     // cells per unit
  int64_t cells_per_piece = 0;
  std::vector<int64_t> piece_begin; // [n_steps + 1]: piece i = cells [piece_begin[i], piece_begin[i + 1]), uploaded in this order
  std::vector<int32_t> units;       // unit ids grouped by the step whose upload makes them computable, ascending within a step
  std::vector<int64_t> step_begin;  // [n_steps + 1]: units of step i = units[step_begin[i] .. step_begin[i + 1])
  int64_t n_late = 0;               // partitioned operators: units with a ghost neighbour; they are not in `units` - the caller applies
                                    // them behind the ghost import, after the last upload
};

// nb: [n_owned][6] neighbour cell (or < 0 on the boundary); unit u = cells [u * unit, (u + 1) * unit)
// allow_ghosts: units with a neighbour >= n_owned are counted in n_late and left out of the steps (false: such a mesh has no plan)
inline HostStreamPlan build_host_stream_plan(const int32_t * nb, int64_t n_owned, int unit, int64_t cells_per_piece, bool allow_ghosts = false)
{
  HostStreamPlan P;
  if (n_owned <= 0 || unit <= 0 || cells_per_piece <= 0) return P;
  const int K = (int)((n_owned + cells_per_piece - 1) / cells_per_piece);
  const int64_t n_units = (n_owned + unit - 1) / unit;
  constexpr int32_t LATE = INT32_MAX;
  std::vector<int32_t> ready((size_t)n_units, 0);
  for (int64_t c = 0; c < n_owned; ++c) {
    int32_t r = (int32_t)(c / cells_per_piece);
    for (int f = 0; f < 6; ++f) {
      const int32_t p = nb[c * 6 + f];
      if (p < 0) continue;
      if (p >= n_owned) { if (!allow_ghosts) return P; r = LATE; continue; } // ghost cell
      r = std::max(r, (int32_t)(p / cells_per_piece));
    }
    int32_t & ru = ready[(size_t)(c / unit)];
    ru = std::max(ru, r);
  }
  for (int64_t u = 0; u < n_units; ++u) if (ready[(size_t)u] == LATE) ++P.n_late;
  P.n_steps = K; P.unit = unit; P.cells_per_piece = cells_per_piece;
  P.piece_begin.resize((size_t)K + 1);
  for (int i = 0; i <= K; ++i) P.piece_begin[i] = std::min<int64_t>(n_owned, (int64_t)i * cells_per_piece);
  P.step_begin.assign((size_t)K + 1, 0);
  for (int64_t u = 0; u < n_units; ++u) if (ready[(size_t)u] != LATE) ++P.step_begin[(size_t)ready[(size_t)u] + 1];
  for (int i = 0; i < K; ++i) P.step_begin[i + 1] += P.step_begin[i];
  P.units.resize((size_t)(n_units - P.n_late));
  std::vector<int64_t> fill(P.step_begin.begin(), P.step_begin.end() - 1);
  for (int64_t u = 0; u < n_units; ++u) if (ready[(size_t)u] != LATE) P.units[(size_t)fill[(size_t)ready[(size_t)u]]++] = (int32_t)u;
  return P;
}

// model of the stream plan, same unit as host_pipeline_model: uploads back to back, the results of step i leave at the same rate as
// soon as upload i is complete and the results of the earlier steps have left (compute time neglected); 1.0 = perfect overlap
inline double host_stream_model(const HostStreamPlan & P)
{
  if (P.n_steps == 0) return 2.0;
  const double n = (double)P.piece_begin[P.n_steps];
  const int64_t n_units = ((int64_t)n + P.unit - 1) / P.unit;
  double t = 0;
  for (int i = 0; i < P.n_steps; ++i) {
    double cells = 0;
    for (int64_t j = P.step_begin[i]; j < P.step_begin[i + 1]; ++j) {
      const int64_t u = P.units[(size_t)j];
      cells += (double)((u + 1 == n_units) ? (int64_t)n - u * P.unit : P.unit);
    }
    t = std::max(t, (double)P.piece_begin[i + 1] / n) + cells / n;
  }
  return t;
}
} // namespace exadg_b200
