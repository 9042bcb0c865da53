// mcw_track.cuh -- the tracking kernel: persistent warps with lane refill, a straight-line
// exit-face search on register-resident element records, parked reflections, and cross sections
// for every material already waiting in the history's own sectors.
//
// Reference behaviour reproduced bit for bit (same IEEE operations in the same order, no FMA):
//   RT TraceRay.C:292-493 exitsElem, RT TraceRayTools.C:371-553 intersectTriangle / intersectQuad,
//   :60-128 lineLineIntersect2D, RT TraceRay.C:1420-1466 neighbour step, :725-782 + ReflectRayBC.C:68-76
//   boundary conditions, MC CollisionKernel.C:143-175 the segment part of onSegment.
//
// Why it is shaped like this (profiles/ncu_r01b_track.txt): the previous version issued ~1000
// thread instructions per segment at 18.5 of 32 lanes and 41 % issue utilisation with 12 warps per SM.
//   * A lane only ever leaves for a collision, a death or the rare general walk: the macroscopic
//     cross sections of EVERY material at the history's energy were computed by the kernel that set
//     that energy (mcw_lookup_all_slots), so entering another material is one 32-byte read.
//   * Specular reflection (two square roots, six divisions) used to run inline with ~1.2 active
//     lanes and cost 17 % of all issued instructions on the reflective pin cell; now the lane parks
//     and the warp reflects all parked lanes together once MCW_REFLECT_MIN of them wait.
//   * The neighbour code carries the side of the neighbour that faces back, so the incoming side
//     needs no search; the source / URR seeds are not re-rounded (idempotent); the trace hash is
//     only kept when asked for; one candidate face is kept instead of four evaluations.
//   * Lanes refill from the queue as soon as MCW_REFILL_MIN are empty, and when the queue has run
//     dry a warp that is down to a few lanes hands its stragglers to the next round instead of
//     walking them at 1/8 efficiency.
#pragma once
#include "mcw_physics.cuh"

#ifndef MCW_TRACK_BLOCK
#define MCW_TRACK_BLOCK 128
#endif
#ifndef MCW_TRACK_MIN_BLOCKS
#define MCW_TRACK_MIN_BLOCKS 4 // register budget 128 per thread -> 16 warps per SM
#endif
#ifndef MCW_HEX_MIN_BLOCKS
#define MCW_HEX_MIN_BLOCKS 3 // HEX8: 24 vertex coordinates and twelve triangles per element; 168 registers
#endif
#ifndef MCW_FUSED_MIN_BLOCKS
#define MCW_FUSED_MIN_BLOCKS 2 // fused (tail) instantiation: 204 registers, 8 warps per SM
#endif
#ifndef MCW_REFILL_MIN
#define MCW_REFILL_MIN 4
#endif
#ifndef MCW_REFLECT_MIN
#define MCW_REFLECT_MIN 4
#endif
#define MCW_DRAIN_MIN_VISIT 32 // a straggler is only handed on after this many segments of progress
#define MCW_FULL 0xffffffffu

// ------------------------------------------------------------------ element record in registers
template <int TYPE>
struct ElemData;

template <>
struct ElemData<MCW_TET4>
{
  V3 v[4];
  int nbr[4]; // side codes (mcw_types.h)
  double hmax;
  int slot, block;
  MCW_D void load(const McwMeshDev& m, int e)
  {
    const double2* q = reinterpret_cast<const double2*>(m.rec + (size_t)e * 128);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), f = __ldg(q + 4), g = __ldg(q + 5);
    int4 n = __ldg(reinterpret_cast<const int4*>(q + 6));
    int4 t = __ldg(reinterpret_cast<const int4*>(q + 7));
    v[0] = v3(a.x, a.y, b.x);
    v[1] = v3(b.y, c.x, c.y);
    v[2] = v3(d.x, d.y, f.x);
    v[3] = v3(f.y, g.x, g.y);
    nbr[0] = n.x, nbr[1] = n.y, nbr[2] = n.z, nbr[3] = n.w;
    hmax = __hiloint2double(t.y, t.x);
    slot = t.z;
    block = t.w;
  }
};

template <>
struct ElemData<MCW_QUAD4>
{
  V3 v[4];
  int nbr[4];
  double hmax;
  int slot, block;
  MCW_D void load(const McwMeshDev& m, int e)
  {
    const double2* q = reinterpret_cast<const double2*>(m.rec + (size_t)e * 128);
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    int4 n = __ldg(reinterpret_cast<const int4*>(q + 4));
    int4 t = __ldg(reinterpret_cast<const int4*>(q + 5));
    v[0] = v3(a.x, a.y, 0.0);
    v[1] = v3(b.x, b.y, 0.0);
    v[2] = v3(c.x, c.y, 0.0);
    v[3] = v3(d.x, d.y, 0.0);
    nbr[0] = n.x, nbr[1] = n.y, nbr[2] = n.z, nbr[3] = n.w;
    hmax = __hiloint2double(t.y, t.x);
    slot = t.z;
    block = t.w;
  }
};

template <>
struct ElemData<MCW_HEX8>
{
  V3 v[8];
  int nbr[6];
  double hmax;
  int slot, block;
  MCW_D void load(const McwMeshDev& m, int e)
  {
    const double2* q = reinterpret_cast<const double2*>(m.rec + (size_t)e * 256);
    double2 r[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
      r[k] = __ldg(q + k);
    int4 n = __ldg(reinterpret_cast<const int4*>(q + 12)); // bytes 192..207: side codes 0-3
    int4 t = __ldg(reinterpret_cast<const int4*>(q + 13)); // 208..223: side codes 4,5, hmax
    int2 w = __ldg(reinterpret_cast<const int2*>(q + 14)); // 224..231: slot, block
    const double* c = reinterpret_cast<const double*>(r);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = v3(c[3 * k], c[3 * k + 1], c[3 * k + 2]);
    nbr[0] = n.x, nbr[1] = n.y, nbr[2] = n.z, nbr[3] = n.w, nbr[4] = t.x, nbr[5] = t.y;
    hmax = __hiloint2double(t.w, t.z);
    slot = w.x;
    block = w.y;
  }
};

template <int N>
MCW_D int mcw_pick(const int (&a)[N], int k)
{
  int r = a[0];
#pragma unroll
  for (int j = 1; j < N; ++j)
    r = (k == j) ? a[j] : r;
  return r;
}

// RT TraceRayTools.C:371-478, the part before the division: the reference's own early-outs plus the
// numerator of `possible`.  Returns false exactly when intersectTriangle would return false without
// looking at the quotient, or when the quotient is certainly <= 0 (numerator <= 0, det >= 1e-8).
struct TriEval
{
  double det, u, v, num;
};
MCW_D bool mcw_tri_stage1(V3 start, V3 dir, V3 p0, V3 p1, V3 p2, double inv_hmax, TriEval& t)
{
  V3 edge1 = (p1 - p0) * inv_hmax;
  V3 edge2 = (p2 - p0) * inv_hmax;
  V3 pvec = cross(dir, edge2);
  t.det = dot(edge1, pvec);
  V3 tvec = (start - p0) * inv_hmax;
  t.u = dot(tvec, pvec);
  V3 qvec = cross(tvec, edge1);
  t.v = dot(dir, qvec);
  t.num = dot(edge2, qvec);
  bool rej = (t.det < MCW_TRACE_TOL) | (t.u < -MCW_TRACE_TOL) | (t.u > t.det + MCW_TRACE_TOL) | (t.v < -MCW_TRACE_TOL) |
             (t.u + t.v > t.det + MCW_TRACE_TOL);
  return !rej && t.num > 0.0;
}

// Both triangles of a quadrilateral side, intersectQuad's (v00, v10, v11) and (v11, v01, v00)
// (RT TraceRayTools.C:514-550), with the second one reusing what is an exact negation of the first:
// its edge2 = (v00 - v11) / hmax = -edge2 of the first, hence pvec' = -pvec, det' = -(edge1' . pvec),
// u' = -(tvec' . pvec), num' = -(edge2 . qvec').  IEEE subtraction, multiplication and the
// left-to-right sums of dot / cross are sign-symmetric under round-to-nearest, so every value is
// bit-identical to evaluating the second triangle from scratch (41 instead of 64 FP64 operations).
MCW_D void mcw_quad_stage1(V3 start, V3 dir, V3 v00, V3 v10, V3 v11, V3 v01, double inv_hmax, TriEval& t1, bool& ok1,
                           TriEval& t2, bool& ok2)
{
  const V3 edge1 = (v10 - v00) * inv_hmax;
  const V3 edge2 = (v11 - v00) * inv_hmax;
  const V3 pvec = cross(dir, edge2);
  t1.det = dot(edge1, pvec);
  const V3 tvec = (start - v00) * inv_hmax;
  t1.u = dot(tvec, pvec);
  const V3 qvec = cross(tvec, edge1);
  t1.v = dot(dir, qvec);
  t1.num = dot(edge2, qvec);
  const bool rej1 = (t1.det < MCW_TRACE_TOL) | (t1.u < -MCW_TRACE_TOL) | (t1.u > t1.det + MCW_TRACE_TOL) |
                    (t1.v < -MCW_TRACE_TOL) | (t1.u + t1.v > t1.det + MCW_TRACE_TOL);
  ok1 = !rej1 && t1.num > 0.0;

  const V3 edge1b = (v01 - v11) * inv_hmax;
  t2.det = -dot(edge1b, pvec);
  const V3 tvecb = (start - v11) * inv_hmax;
  t2.u = -dot(tvecb, pvec);
  const V3 qvecb = cross(tvecb, edge1b);
  t2.v = dot(dir, qvecb);
  t2.num = -dot(edge2, qvecb);
  const bool rej2 = (t2.det < MCW_TRACE_TOL) | (t2.u < -MCW_TRACE_TOL) | (t2.u > t2.det + MCW_TRACE_TOL) |
                    (t2.v < -MCW_TRACE_TOL) | (t2.u + t2.v > t2.det + MCW_TRACE_TOL);
  ok2 = !rej2 && t2.num > 0.0;
}

// true when the hit is classified as a vertex or an edge (RT TraceRayTools.C:436-474); the general
// kernel then redoes the step, so the fast path does not need to know which one
MCW_D bool mcw_tri_at_extrema(const TriEval& t)
{
  const double lo = MCW_TRACE_TOL * t.det;
  return (t.u < lo) | (t.v < lo) | (t.u + t.v > (1.0 - MCW_TRACE_TOL) * t.det);
}

// vertex / edge classification of a triangle hit (RT TraceRayTools.C:436-474)
MCW_D void mcw_tri_classify(const TriEval& t, int v0, int v1, int v2, Extrema& ext)
{
  const double u = t.u, v = t.v, det = t.det;
  if (u < MCW_TRACE_TOL * det)
  {
    if (v < MCW_TRACE_TOL * det)
      ext.setVertex(v0);
    else if (v > (1.0 - MCW_TRACE_TOL) * det)
      ext.setVertex(v2);
    else
      ext.setEdge(v0, v2);
  }
  else if (v < MCW_TRACE_TOL * det)
  {
    if (u > (1.0 - MCW_TRACE_TOL) * det)
      ext.setVertex(v1);
    else
      ext.setEdge(v0, v1);
  }
  else if (u + v > (1.0 - MCW_TRACE_TOL) * det)
    ext.setEdge(v1, v2);
}

// Exit search.  Returns true with (iside, idist, ipoint) when exactly the plain case applies: one
// side is hit away from its vertices and edges.  Returns false when no side is hit or the hit is at
// a vertex / edge -- the general kernel (mcw_fallback.cuh) then redoes the step from unchanged
// state with mcw_exits_elem<TYPE> (mcw_device.cuh), which stays the definition of the search.
template <int TYPE>
MCW_D bool mcw_exit_fast(const McwMeshDev& mesh, const ElemData<TYPE>& D, double inv_hmax, int cur, int incoming_side,
                         V3 start, V3 dir, V3& ipoint, int& iside, double& idist);

template <>
MCW_D bool mcw_exit_fast<MCW_TET4>(const McwMeshDev& mesh, const ElemData<MCW_TET4>& D, double inv_hmax, int cur, int incoming_side,
                                   V3 start, V3 dir, V3& ipoint, int& iside, double& idist)
{
  const double hmax = D.hmax; // inv_hmax = 1.0 / hmax, computed by the caller when hmax changes
  TriEval keep;
  keep.det = 1.0;
  keep.u = keep.v = keep.num = 0.0;
  int n_cand = 0;
  iside = MCW_INVALID;
#pragma unroll
  for (int s = 0; s < 4; ++s)
  {
    const int a = side_node<MCW_TET4>(s, 2), b = side_node<MCW_TET4>(s, 1), c = side_node<MCW_TET4>(s, 0);
    TriEval t;
    const bool ok = mcw_tri_stage1(start, dir, D.v[a], D.v[b], D.v[c], inv_hmax, t) && (s != incoming_side);
    if (ok)
    {
      keep = t;
      iside = s;
      ++n_cand;
    }
  }
  if (n_cand != 1)
  {
    if (n_cand == 0)
      return false;
    // two faces pass the rejection tests (the ray leaves through their common edge or vertex, or
    // grazes a tolerance band): the reference routine decides, farthest hit wins
    ElemRec<MCW_TET4> R(mesh, cur);
    Extrema ext;
    if (!mcw_exits_elem<MCW_TET4>(mesh, R, incoming_side, start, dir, ipoint, iside, ext, idist))
      return false;
    return !ext.atExtrema();
  }
  const double possible = keep.num / keep.det;
  const double cd = possible * hmax;
  // RT TraceRayTools.C:428 `possible <= tol` and RT TraceRay.C:435-444 `distance > 1e-8 * hmax`
  if (possible <= MCW_TRACE_TOL || !(cd > MCW_TRACE_TOL * hmax))
    return false;
  idist = cd;
  ipoint = start + idist * dir;
  return !mcw_tri_at_extrema(keep);
}

template <>
MCW_D bool mcw_exit_fast<MCW_HEX8>(const McwMeshDev& mesh, const ElemData<MCW_HEX8>& D, double inv_hmax, int cur, int incoming_side,
                                   V3 start, V3 dir, V3& ipoint, int& iside, double& idist)
{
  // RT TraceRayTools.C:480-553 intersectQuad(v00, v10, v11, v01): triangle (v00, v10, v11) first,
  // then (v11, v01, v00) if the first did not hit.
  const double hmax = D.hmax;
  unsigned c1 = 0, c2 = 0;
  TriEval keep;
  keep.det = 1.0;
  keep.u = keep.v = keep.num = 0.0;
#pragma unroll
  for (int s = 0; s < 6; ++s)
  {
    const int v00 = side_node<MCW_HEX8>(s, 3), v10 = side_node<MCW_HEX8>(s, 2), v11 = side_node<MCW_HEX8>(s, 1),
              v01 = side_node<MCW_HEX8>(s, 0);
    TriEval t1, t2;
    bool ok1, ok2;
    mcw_quad_stage1(start, dir, D.v[v00], D.v[v10], D.v[v11], D.v[v01], inv_hmax, t1, ok1, t2, ok2);
    const bool live = s != incoming_side;
    if (ok1 && live)
    {
      c1 |= 1u << s;
      keep = t1;
    }
    if (ok2 && live)
    {
      c2 |= 1u << s;
      if (!ok1)
        keep = t2;
    }
  }
  const unsigned sides = c1 | c2;
  iside = MCW_INVALID;
  if (sides == 0)
    return false;
  if ((sides & (sides - 1)) == 0 && (c1 & c2) == 0)
  {
    // exactly one side and one triangle can hit: `keep` holds its evaluation
    const int s = __ffs(sides) - 1;
    const double possible = keep.num / keep.det;
    const double cd = possible * hmax;
    if (possible <= MCW_TRACE_TOL || !(cd > MCW_TRACE_TOL * hmax))
      return false;
    iside = s;
    idist = cd;
    ipoint = start + idist * dir;
    const int v00 = side_node<MCW_HEX8>(s, 3), v10 = side_node<MCW_HEX8>(s, 2), v11 = side_node<MCW_HEX8>(s, 1),
              v01 = side_node<MCW_HEX8>(s, 0);
    Extrema ext;
    if (c1)
      mcw_tri_classify(keep, v00, v10, v11, ext);
    else
      mcw_tri_classify(keep, v11, v01, v00, ext);
    if (ext.atEdgeOf(v00, v11))
      ext.invalidate(); // the quad's diagonal is not an edge (RT TraceRayTools.C:545-550)
    return !ext.atExtrema();
  }
  // several candidate sides, or both triangles of one side: the reference routine decides
  ElemRec<MCW_HEX8> R(mesh, cur);
  Extrema ext;
  if (!mcw_exits_elem<MCW_HEX8>(mesh, R, incoming_side, start, dir, ipoint, iside, ext, idist))
    return false;
  return !ext.atExtrema();
}

template <>
MCW_D bool mcw_exit_fast<MCW_QUAD4>(const McwMeshDev& mesh, const ElemData<MCW_QUAD4>& D, double inv_hmax, int cur, int incoming_side,
                                    V3 start, V3 dir, V3& ipoint, int& iside, double& idist)
{
  // RT TraceRayTools.C:60-128 lineLineIntersect2D per side.  Stage 1 rejects, without dividing, the
  // sides whose parameters t or u are outside their windows by more than twice the tolerance (the
  // quotients' rounding error is 1e-16 relative, the margin 1e-8 absolute); stage 2 is the
  // reference arithmetic on the survivors.
  (void)inv_hmax;
  const double length = mesh.domain_max_length;
  const V3 r = dir * length;
  const double M = 2.0 * MCW_TRACE_TOL;
  unsigned surv = 0;
  double rxs_[4], tn_[4], un_[4];
#pragma unroll
  for (int s = 0; s < 4; ++s)
  {
    const V3 v0 = D.v[side_node<MCW_QUAD4>(s, 0)], v1 = D.v[side_node<MCW_QUAD4>(s, 1)];
    const V3 sv = v1 - v0;
    const double rxs = r.x * sv.y - r.y * sv.x;
    const V3 d0 = v0 - start;
    const double tn = d0.x * sv.y - d0.y * sv.x;
    const double un = d0.x * r.y - d0.y * r.x;
    rxs_[s] = rxs, tn_[s] = tn, un_[s] = un;
    const double ar = fabs(rxs);
    const double ts = rxs < 0 ? -tn : tn, us = rxs < 0 ? -un : un; // numerators with the sign of the quotient
    // certainly rejected by the reference: |rxs| < tol, t <= -tol or t > 1 + tol, u <= -tol or u > 1 + tol
    bool rej = (ar < MCW_TRACE_TOL) | (ts < -M * ar) | (ts > (1.0 + M) * ar) | (us < -M * ar) | (us > (1.0 + M) * ar);
    surv |= (!rej && s != incoming_side) ? (1u << s) : 0u;
  }
  double best = MCW_TRACE_TOL * D.hmax;
  iside = MCW_INVALID;
  bool at_vertex = false;
  while (surv)
  {
    const int s = __ffs(surv) - 1;
    surv &= surv - 1;
    double rxs = rxs_[0], tn = tn_[0], un = un_[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (s == j)
        rxs = rxs_[j], tn = tn_[j], un = un_[j];
    const double t = tn / rxs;
    if (0 >= t + MCW_TRACE_TOL || t - MCW_TRACE_TOL > 1.0)
      continue;
    const double u = un / rxs;
    if (!(0 < u + MCW_TRACE_TOL && u - MCW_TRACE_TOL <= 1.0))
      continue;
    const double cd = t * length;
    if (cd > best)
    {
      iside = s;
      idist = cd;
      best = cd;
      ipoint = start + r * t;
      at_vertex = (u < MCW_TRACE_TOL) | (u > 1.0 - MCW_TRACE_TOL);
    }
  }
  return iside != MCW_INVALID && !at_vertex;
}

#ifndef MCW_FUSED_PREFETCH
#define MCW_FUSED_PREFETCH 0
#endif
MCW_D void mcw_prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------ the kernel
// One lane = one history in flight.  Lane state lives in registers between segments and is written
// back only when the history leaves the kernel.
enum
{
  MCW_LN_EMPTY = 0,  // no history: waits for a refill (or for the end of the launch)
  MCW_LN_TRACK = 1,  // walks one element per iteration
  MCW_LN_REFLECT = 2, // parked on a reflecting boundary side (inc_side = that side)
  MCW_LN_COLLIDE = 3  // fused mode only: parked at a collision point, state written back to the SoA
};
// fused mode: parked collisions are processed together once `collide_min` of them wait (kernel argument)

// FUSED = false: the event-based kernel (a lane leaves at its collision; k_collide runs afterwards on
// the whole queue).  FUSED = true: the tail of a generation, when fewer histories are left than the
// device has lanes -- rounds would be as long as the slowest flight while most lanes idle, so each
// lane keeps its history through its collisions (mcw_collide_history, batched per warp) until it
// dies.  Only the rare vertex / edge steps still leave through the fallback queue.
template <int TYPE, bool FUSED>
__global__ void __launch_bounds__(MCW_TRACK_BLOCK, (FUSED ? MCW_FUSED_MIN_BLOCKS : (TYPE == MCW_HEX8 ? MCW_HEX_MIN_BLOCKS : MCW_TRACK_MIN_BLOCKS)))
k_track(McwParams P, const int32_t* __restrict__ queue, const unsigned long long* __restrict__ count,
        unsigned long long* __restrict__ head, int32_t* __restrict__ q_fly_next, unsigned long long* __restrict__ cnt_next,
        int32_t* __restrict__ q_collide, int32_t* __restrict__ q_fallback, int drain_lanes, int collide_min)
{
  const long long n_q = (long long)*count;
  const unsigned lane = threadIdx.x & 31u;
  McwParticlesDev& p = P.p;
  const bool track_tally = P.tallies_active && P.any_tracklength;
  const bool keep_hash = P.record_traces != 0;

  int state = MCW_LN_EMPTY;
  bool exhausted = n_q == 0;
  int i = -1;
  V3 point = v3(0, 0, 0), dir = v3(0, 0, 1);
  int cur = 0, last = -1, inc_side = -1, slot = -1, n_seg = 0, visit = 0;
  SegState s;
  s.E = s.wgt = s.sig_a = s.sig_f = s.sig_nf = s.k_trk = 0.0;
  s.sig_t = 1.0;
  s.seed_trk = s.seed_src = s.seed_urr = 0;
  uint64_t hash = 0;
  unsigned segs = 0;
  double hmax_c = 0.0, inv_hmax = 0.0; // 1.0 / hmax, recomputed only when hmax (a block constant) changes

  auto load_lane = [&]() {
    point = v3(p.x[i], p.y[i], p.z[i]);
    dir = v3(p.ux[i], p.uy[i], p.uz[i]);
    cur = p.elem[i];
    last = p.last_elem[i];
    inc_side = p.inc_side[i];
    s.E = p.E[i];
    s.wgt = p.wgt[i];
    s.seed_trk = p.seed_trk[i];
    if (MCW_RARE && P.lib.n_urr)
      s.seed_urr = p.seed_urr[i]; // track-length scores of a nuclide in its unresolved range need the band
    s.k_trk = p.k_trk[i];
    n_seg = p.n_seg[i];
    if (keep_hash)
      hash = p.hash[i];
    slot = -1; // the cross sections of the current material are (re)read from the history's sectors
#if MCW_FUSED_PREFETCH > 1
    if (FUSED)
      mcw_prefetch_l1(p.xs + (size_t)i * P.n_slots * 4);
#endif
  };

  for (;;)
  {
    const unsigned m_empty = __ballot_sync(MCW_FULL, state == MCW_LN_EMPTY);
    if (m_empty == MCW_FULL && exhausted)
      break;
    if (!exhausted && __popc(m_empty) >= MCW_REFILL_MIN)
    {
      const int n_idle = __popc(m_empty);
      unsigned long long base = 0;
      if (lane == 0)
        base = atomicAdd(head, (unsigned long long)n_idle);
      base = __shfl_sync(MCW_FULL, base, 0);
      if ((long long)base + n_idle >= n_q)
        exhausted = true;
      const long long mine = (long long)base + __popc(m_empty & ((1u << lane) - 1u));
      if (state == MCW_LN_EMPTY && mine < n_q)
      {
        i = queue[mine];
        const int st = p.status[i];
        const int ef = p.ext_first[i];
        if (st == MCW_ST_FLY && ef != MCW_INVALID)
        {
          // last hit was a vertex/edge: general walk (RT TraceRay.C:1134)
          p.status[i] = MCW_ST_FALLBACK;
          q_fallback[mcw_warp_append(&P.counters->n_fallback, true)] = i;
        }
        else if (st == MCW_ST_FLY)
        {
          load_lane();
          visit = 0;
          state = MCW_LN_TRACK;
        }
      }
    }
    // a warp that can no longer refill and is down to a few histories hands them to the next round
    const bool drain = exhausted && (32 - __popc(__ballot_sync(MCW_FULL, state == MCW_LN_EMPTY))) < drain_lanes;

    const unsigned m_refl = __ballot_sync(MCW_FULL, state == MCW_LN_REFLECT);
    if (m_refl != 0u)
    {
      const unsigned m_track = __ballot_sync(MCW_FULL, state == MCW_LN_TRACK);
      if (__popc(m_refl) >= MCW_REFLECT_MIN || m_track == 0u || drain)
      {
        if (state == MCW_LN_REFLECT)
        {
          // RT ReflectRayBC.C:68-76 + Ray::changeDirection (Ray.C:206-214); inc_side holds the side hit
          ElemRec<TYPE> R(P.mesh, cur);
          V3 n = mcw_side_normal<TYPE>(R, inc_side);
          double dn = dot(dir, n);
          V3 r = dir - (2.0 * dn) * n;
          r = r / norm(r);
          dir = unit(r);
          last = cur;
          state = MCW_LN_TRACK;
        }
      }
    }

    if (FUSED)
    {
      const unsigned m_coll = __ballot_sync(MCW_FULL, state == MCW_LN_COLLIDE);
      if (m_coll != 0u)
      {
        const unsigned m_track = __ballot_sync(MCW_FULL, state == MCW_LN_TRACK);
        if (__popc(m_coll) >= collide_min || m_track == 0u)
        {
          // all 32 lanes enter (full-warp bank append); only the parked ones collide
          const bool mine = state == MCW_LN_COLLIDE;
          const bool survived = mcw_collide_history<TYPE, false>(P, mine ? i : 0, mine) && mine;
          // cross sections at the survivors' new energies: the warp shares each history's nuclide lookups
          mcw_warp_lookup_all_slots(P, __ballot_sync(MCW_FULL, survived), i, survived ? p.E[i] : 0.0);
          if (mine)
          {
            if (survived)
            {
              load_lane();
              state = MCW_LN_TRACK;
            }
            else
              state = MCW_LN_EMPTY;
          }
        }
      }
    }

    if (state == MCW_LN_TRACK)
    {
      int status = MCW_ST_FLY, fate = -1;
      bool leave = false, requeue = false;
      ElemData<TYPE> D;
      D.load(P.mesh, cur);
#if MCW_FUSED_PREFETCH
      if (FUSED)
      {
        // the tail is bound by the latency of single histories (few live lanes per warp, DRAM and L2 nearly idle):
        // pull the records of all neighbours into L1 while the exit search runs, so the next element's record
        // is an L1 hit instead of an L2 round trip on the history's critical path
#pragma unroll
        for (int sd = 0; sd < (int)(sizeof(D.nbr) / sizeof(int)); ++sd)
          if (D.nbr[sd] >= 0 && sd != inc_side)
            mcw_prefetch_l1(P.mesh.rec + (size_t)(D.nbr[sd] >> MCW_NBR_SHIFT) * (TYPE == MCW_HEX8 ? 256 : 128));
      }
#endif
      // the segment's random number and logarithm, ahead of the exit search (independent work while
      // the record arrives); the seed is committed only if the segment is taken
      uint64_t seed_next = s.seed_trk;
      const double neg_log = -mcw_log(mcw_prn(seed_next));
      if (D.hmax != hmax_c)
      {
        hmax_c = D.hmax;
        inv_hmax = 1.0 / hmax_c;
      }
      if (D.slot != slot)
      {
        // another material at the same energy: its cross sections are already in this history's sector
        slot = D.slot;
        const double2* c = reinterpret_cast<const double2*>(p.xs + ((size_t)i * P.n_slots + slot) * 4);
        const double2 a = c[0], b = c[1];
        s.sig_t = a.x;
        s.sig_nf = b.y;
        if (track_tally)
        {
          s.sig_a = a.y;
          s.sig_f = b.x;
        }
      }
      if (n_seg >= P.max_events)
      {
        status = MCW_ST_DEAD;
        fate = 3;
        atomicAdd(&P.counters->n_failed, 1ull);
        leave = true;
      }
      else if (visit >= P.visit_cap || (drain && visit >= MCW_DRAIN_MIN_VISIT))
      {
        requeue = true;
        leave = true;
      }
      else
      {
        V3 ipoint;
        int iside;
        double idist;
        if (!mcw_exit_fast<TYPE>(P.mesh, D, inv_hmax, cur, inc_side, point, dir, ipoint, iside, idist))
        {
          status = MCW_ST_FALLBACK; // nothing modified: the general kernel redoes this step
          leave = true;
        }
        else
        {
          ++n_seg;
          ++segs;
          ++visit;
          V3 pos;
          const int code = mcw_pick(D.nbr, iside);
          const int block = D.block;
          const int elem = cur;
          s.seed_trk = seed_next;
          const bool collided = mcw_segment_nl(P, s, neg_log, elem, block, P.slot_mat[slot], ipoint, dir, idist, pos);
          if (keep_hash)
            hash = mcw_hash_step(hash, elem, collided ? 15 : iside);
          if (collided)
          {
            point = pos;
            status = MCW_ST_COLLIDE;
            leave = true;
            if (MCW_RARE && P.trace)
              p.tx[i] = ipoint.x, p.ty[i] = ipoint.y, p.tz[i] = ipoint.z;
          }
          else
          {
            point = ipoint;
            if (code >= 0)
            {
              // RT TraceRay.C:1433-1458: neighbour and the side of it that faces back
              last = cur;
              cur = code >> MCW_NBR_SHIFT;
              inc_side = code & ((1 << MCW_NBR_SHIFT) - 1);
            }
            else
            {
              // RT TraceRay.C:725-782 external boundary, :2043 onBoundary
              const int bid = -code - 1;
              const int kind = bid < P.mesh.n_bnd ? __ldg(P.mesh.bnd_kind + bid) : MCW_BC_KILL;
              inc_side = iside;
              if (kind == MCW_BC_KILL)
              {
                status = MCW_ST_DEAD;
                fate = 1;
                leave = true;
              }
              else
                state = MCW_LN_REFLECT;
            }
            if (MCW_RARE && P.trace)
              mcw_trace_emit(P, i, n_seg, 1, ipoint, s.E, s.wgt, p.n_progeny[i], s.seed_trk, p.seed_src[i],
                             p.seed_urr[i]);
          }
        }
      }
      if (leave)
      {
        p.x[i] = point.x;
        p.y[i] = point.y;
        p.z[i] = point.z;
        p.elem[i] = cur;
        p.seed_trk[i] = s.seed_trk;
        p.k_trk[i] = s.k_trk;
        p.n_seg[i] = n_seg;
        if (keep_hash)
          p.hash[i] = hash;
        p.status[i] = (int8_t)(requeue ? MCW_ST_FLY : status);
        p.ux[i] = dir.x;
        p.uy[i] = dir.y;
        p.uz[i] = dir.z;
        p.last_elem[i] = last;
        p.inc_side[i] = (int8_t)inc_side;
        if (fate >= 0)
          p.fate[i] = (int8_t)fate;
        if (requeue)
        {
          if (!P.ordered_queues)
            q_fly_next[mcw_warp_append(cnt_next, true)] = i;
        }
        else if (status == MCW_ST_COLLIDE)
        {
          if (!FUSED && !P.ordered_queues)
            q_collide[mcw_warp_append(&P.counters->n_collide, true)] = i;
        }
        else if (status == MCW_ST_FALLBACK)
          q_fallback[mcw_warp_append(&P.counters->n_fallback, true)] = i;
        state = (FUSED && !requeue && status == MCW_ST_COLLIDE) ? MCW_LN_COLLIDE : MCW_LN_EMPTY;
      }
    }
  }
  // segment counter: one atomic per warp
  {
    unsigned tot = segs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      tot += __shfl_xor_sync(MCW_FULL, tot, o);
    if (lane == 0 && tot)
      atomicAdd(&P.counters->n_segments, (unsigned long long)tot);
  }
}
