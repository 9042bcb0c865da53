// mcw_track.cuh -- the tracking kernel: persistent warps with lane refill, a straight-line
// exit-face search on register-resident element records, and a per-particle macroscopic
// cross-section cache so a material change does not end the visit.
//
// Reference behaviour reproduced bit for bit (same IEEE operations in the same order, no FMA):
//   RT TraceRay.C:292-493 exitsElem, RT TraceRayTools.C:371-553 intersectTriangle / intersectQuad,
//   :60-128 lineLineIntersect2D, RT TraceRay.C:1420-1466 neighbour step, :725-782 + ReflectRayBC.C:68-76
//   boundary conditions, MC CollisionKernel.C:143-175 the segment part of onSegment.
//
// Why it is shaped like this (profiles/track_r01a_baseline.md): with one thread per history and a
// loop "until collision", lanes of a warp finish after wildly different segment counts and the
// baseline kernel ran with 4.3 of 32 lanes active.  Here a warp keeps all lanes busy: a lane whose
// history stops (collision, unknown cross section, boundary kill, rare-path hand-off) parks, and
// when MCW_REFILL_MIN lanes are parked the warp pulls that many new histories from the queue with
// one atomicAdd.  The exit search is split in two exact stages: stage 1 evaluates, for every side
// at once and without divisions, the reference's own rejection tests (det, u, v) plus the sign of
// the distance numerator; stage 2 runs the reference arithmetic (division, tolerance and
// farthest-hit selection) only for the sides that survived -- normally one.
#pragma once
#include "mcw_physics.cuh"

#define MCW_TRACK_BLOCK 128
#define MCW_REFILL_MIN 8
#define MCW_FULL 0xffffffffu

// ------------------------------------------------------------------ element record in registers
template <int TYPE>
struct ElemData;

template <>
struct ElemData<MCW_TET4>
{
  V3 v[4];
  int nbr[4];
  double hmax;
  int material, block;
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
    material = t.z;
    block = t.w;
  }
};

template <>
struct ElemData<MCW_QUAD4>
{
  V3 v[4];
  int nbr[4];
  double hmax;
  int material, block;
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
    material = t.z;
    block = t.w;
  }
};

template <>
struct ElemData<MCW_HEX8>
{
  V3 v[8];
  int nbr[6];
  double hmax;
  int material, block;
  MCW_D void load(const McwMeshDev& m, int e)
  {
    const double2* q = reinterpret_cast<const double2*>(m.rec + (size_t)e * 256);
    double2 r[12];
#pragma unroll
    for (int k = 0; k < 12; ++k)
      r[k] = __ldg(q + k);
    int4 n = __ldg(reinterpret_cast<const int4*>(q + 12)); // bytes 192..207: nbr 0-3
    int4 t = __ldg(reinterpret_cast<const int4*>(q + 13)); // 208..223: nbr 4,5, hmax
    int2 w = __ldg(reinterpret_cast<const int2*>(q + 14)); // 224..231: material, block
    const double* c = reinterpret_cast<const double*>(r);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = v3(c[3 * k], c[3 * k + 1], c[3 * k + 2]);
    nbr[0] = n.x, nbr[1] = n.y, nbr[2] = n.z, nbr[3] = n.w, nbr[4] = t.x, nbr[5] = t.y;
    hmax = __hiloint2double(t.w, t.z);
    material = w.x;
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

// Exit search.  Same contract and same results as mcw_exits_elem<TYPE> (mcw_device.cuh), which stays
// the definition: the rare multi-candidate hexahedron case calls it directly.
template <int TYPE>
MCW_D bool mcw_exit_fast(const McwMeshDev& mesh, const ElemData<TYPE>& D, int cur, int incoming_side, V3 start, V3 dir,
                         V3& ipoint, int& iside, Extrema& iext, double& idist);

template <>
MCW_D bool mcw_exit_fast<MCW_TET4>(const McwMeshDev& mesh, const ElemData<MCW_TET4>& D, int cur, int incoming_side,
                                   V3 start, V3 dir, V3& ipoint, int& iside, Extrema& iext, double& idist)
{
  const double hmax = D.hmax;
  const double inv_hmax = 1.0 / hmax;
  TriEval t[4];
  unsigned surv = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s)
  {
    const int a = side_node<MCW_TET4>(s, 2), b = side_node<MCW_TET4>(s, 1), c = side_node<MCW_TET4>(s, 0);
    bool ok = mcw_tri_stage1(start, dir, D.v[a], D.v[b], D.v[c], inv_hmax, t[s]);
    surv |= (ok && s != incoming_side) ? (1u << s) : 0u;
  }
  double best = MCW_TRACE_TOL * hmax;
  iside = MCW_INVALID;
  TriEval tb;
  tb.det = tb.u = tb.v = tb.num = 0.0;
  while (surv)
  {
    const int s = __ffs(surv) - 1;
    surv &= surv - 1;
    TriEval c = t[0];
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (s == j)
        c = t[j];
    const double possible = c.num / c.det;
    if (possible <= MCW_TRACE_TOL)
      continue;
    const double cd = possible * hmax;
    if (cd > best)
    {
      iside = s;
      idist = cd;
      best = cd;
      tb = c;
    }
  }
  if (iside == MCW_INVALID)
    return false;
  ipoint = start + idist * dir;
  iext.invalidate();
  mcw_tri_classify(tb, side_node<MCW_TET4>(iside, 2), side_node<MCW_TET4>(iside, 1), side_node<MCW_TET4>(iside, 0), iext);
  return true;
}

template <>
MCW_D bool mcw_exit_fast<MCW_HEX8>(const McwMeshDev& mesh, const ElemData<MCW_HEX8>& D, int cur, int incoming_side,
                                   V3 start, V3 dir, V3& ipoint, int& iside, Extrema& iext, double& idist)
{
  // RT TraceRayTools.C:480-553 intersectQuad(v00, v10, v11, v01): triangle (v00, v10, v11) first,
  // then (v11, v01, v00) if the first did not hit.
  const double hmax = D.hmax;
  const double inv_hmax = 1.0 / hmax;
  unsigned c1 = 0, c2 = 0;
  TriEval keep;
  keep.det = keep.u = keep.v = keep.num = 0.0;
#pragma unroll
  for (int s = 0; s < 6; ++s)
  {
    const int v00 = side_node<MCW_HEX8>(s, 3), v10 = side_node<MCW_HEX8>(s, 2), v11 = side_node<MCW_HEX8>(s, 1),
              v01 = side_node<MCW_HEX8>(s, 0);
    TriEval t1, t2;
    bool ok1 = mcw_tri_stage1(start, dir, D.v[v00], D.v[v10], D.v[v11], inv_hmax, t1);
    bool ok2 = mcw_tri_stage1(start, dir, D.v[v11], D.v[v01], D.v[v00], inv_hmax, t2);
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
  if (sides == 0)
  {
    iside = MCW_INVALID;
    return false;
  }
  if ((sides & (sides - 1)) == 0 && (c1 & c2) == 0)
  {
    // exactly one side and one triangle can hit: `keep` holds its evaluation
    const int s = __ffs(sides) - 1;
    const double possible = keep.num / keep.det;
    const double cd = possible * hmax;
    if (possible <= MCW_TRACE_TOL || !(cd > MCW_TRACE_TOL * hmax))
    {
      iside = MCW_INVALID;
      return false;
    }
    iside = s;
    idist = cd;
    ipoint = start + idist * dir;
    const int v00 = side_node<MCW_HEX8>(s, 3), v10 = side_node<MCW_HEX8>(s, 2), v11 = side_node<MCW_HEX8>(s, 1),
              v01 = side_node<MCW_HEX8>(s, 0);
    iext.invalidate();
    if (c1)
      mcw_tri_classify(keep, v00, v10, v11, iext);
    else
      mcw_tri_classify(keep, v11, v01, v00, iext);
    if (iext.atEdgeOf(v00, v11))
      iext.invalidate();
    return true;
  }
  // several candidate sides, or both triangles of one side: the reference routine decides
  ElemRec<MCW_HEX8> R(mesh, cur);
  return mcw_exits_elem<MCW_HEX8>(mesh, R, incoming_side, start, dir, ipoint, iside, iext, idist);
}

template <>
MCW_D bool mcw_exit_fast<MCW_QUAD4>(const McwMeshDev& mesh, const ElemData<MCW_QUAD4>& D, int cur, int incoming_side,
                                    V3 start, V3 dir, V3& ipoint, int& iside, Extrema& iext, double& idist)
{
  // RT TraceRayTools.C:60-128 lineLineIntersect2D per side.  Stage 1 rejects, without dividing, the
  // sides whose parameters t or u are outside their windows by more than twice the tolerance (the
  // quotients' rounding error is 1e-16 relative, the margin 1e-8 absolute); stage 2 is the
  // reference arithmetic on the survivors.
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
      iext.invalidate();
      if (u < MCW_TRACE_TOL)
        iext.setVertex(side_node<MCW_QUAD4>(s, 0));
      else if (u > 1.0 - MCW_TRACE_TOL)
        iext.setVertex(side_node<MCW_QUAD4>(s, 1));
    }
  }
  return iside != MCW_INVALID;
}

// ------------------------------------------------------------------ the kernel
// One lane = one history in flight.  Lane state lives in registers between segments and is written
// back only when the history leaves the kernel.
template <int TYPE>
__global__ void __launch_bounds__(MCW_TRACK_BLOCK)
k_track(McwParams P, const int32_t* __restrict__ queue, const unsigned long long* __restrict__ count,
        unsigned long long* __restrict__ head, int32_t* __restrict__ q_fly_next, unsigned long long* __restrict__ cnt_next,
        int32_t* __restrict__ q_collide, int32_t* __restrict__ q_fallback)
{
  using T = McwElem<TYPE>;
  const long long n_q = (long long)*count;
  const unsigned lane = threadIdx.x & 31u;
  McwParticlesDev& p = P.p;

  bool active = false;
  bool exhausted = n_q == 0;
  int i = -1;
  V3 point = v3(0, 0, 0), dir = v3(0, 0, 1);
  int cur = 0, last = -1, inc_side = -1, xs_mat = -1, n_seg = 0, visit = 0;
  bool need_inc = false;
  unsigned valid = 0;
  SegState s;
  s.E = s.wgt = s.sig_t = s.sig_a = s.sig_f = s.sig_nf = s.k_trk = 0.0;
  s.seed_trk = s.seed_src = s.seed_urr = 0;
  uint64_t hash = 0;
  unsigned long long segs = 0;

  for (;;)
  {
    const unsigned idle = __ballot_sync(MCW_FULL, !active);
    if (idle == MCW_FULL && exhausted)
      break;
    if (!exhausted && __popc(idle) >= MCW_REFILL_MIN)
    {
      const int n_idle = __popc(idle);
      unsigned long long base = 0;
      if (lane == 0)
        base = atomicAdd(head, (unsigned long long)n_idle);
      base = __shfl_sync(MCW_FULL, base, 0);
      if ((long long)base + n_idle >= n_q)
        exhausted = true;
      const long long mine = (long long)base + __popc(idle & ((1u << lane) - 1u));
      if (!active && mine < n_q)
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
          point = v3(p.x[i], p.y[i], p.z[i]);
          dir = v3(p.ux[i], p.uy[i], p.uz[i]);
          cur = p.elem[i];
          last = p.last_elem[i];
          inc_side = p.inc_side[i];
          s.E = p.E[i];
          s.wgt = p.wgt[i];
          s.seed_trk = p.seed_trk[i];
          s.seed_src = p.seed_src[i];
          s.seed_urr = p.seed_urr[i];
          s.sig_t = p.sig_t[i];
          s.sig_a = p.sig_a[i];
          s.sig_f = p.sig_f[i];
          s.sig_nf = p.sig_nf[i];
          s.k_trk = p.k_trk[i];
          xs_mat = p.xs_mat[i];
          valid = P.xs_slots ? p.xs_valid[i] : 0u;
          n_seg = p.n_seg[i];
          hash = p.hash[i];
          need_inc = false;
          visit = 0;
          active = true;
        }
      }
    }
    if (active)
    {
      int status = MCW_ST_FLY, fate = -1;
      bool leave = false, requeue = false;
      ElemData<TYPE> D;
      D.load(P.mesh, cur);
      if (need_inc)
      {
        // RT TraceRay.C:1456-1458: the side of the new element that faces the one we came from
        int s_in = MCW_INVALID;
#pragma unroll
        for (int k = T::NS - 1; k >= 0; --k)
          if (D.nbr[k] == last)
            s_in = k;
        inc_side = s_in;
        need_inc = false;
      }
      if (D.material != xs_mat)
      {
        // new material at the same energy: take the cached macroscopic cross sections if this
        // history already looked them up since its last collision, else go back to the xs kernel
        if (P.xs_slots && ((valid >> D.material) & 1u))
        {
          const double* c = p.xs_cache + ((size_t)D.material * 4) * (size_t)P.xs_stride + i;
          s.sig_t = c[0];
          s.sig_a = c[(size_t)P.xs_stride];
          s.sig_f = c[2 * (size_t)P.xs_stride];
          s.sig_nf = c[3 * (size_t)P.xs_stride];
          xs_mat = D.material;
        }
        else
        {
          requeue = true;
          leave = true;
        }
      }
      if (!leave && (n_seg >= P.max_events))
      {
        status = MCW_ST_DEAD;
        fate = 3;
        atomicAdd(&P.counters->n_failed, 1ull);
        leave = true;
      }
      if (!leave && visit >= P.visit_cap)
      {
        requeue = true; // bound the time one lane can hold its warp after the queue ran dry
        leave = true;
      }
      if (!leave)
      {
        V3 ipoint;
        int iside;
        Extrema ext;
        double idist;
        if (!mcw_exit_fast<TYPE>(P.mesh, D, cur, inc_side, point, dir, ipoint, iside, ext, idist) || ext.atExtrema())
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
          const bool collided = mcw_segment(P, s, cur, D.block, xs_mat, ipoint, dir, idist, pos);
          hash = mcw_hash_step(hash, cur, collided ? 15 : iside);
          if (collided)
          {
            point = pos;
            status = MCW_ST_COLLIDE;
            leave = true;
          }
          else
          {
            point = ipoint;
            const int nb = mcw_pick(D.nbr, iside);
            if (nb >= 0)
            {
              last = cur;
              cur = nb;
              need_inc = true;
            }
            else
            {
              // RT TraceRay.C:725-782 external boundary, :2043 onBoundary
              const int bid = -nb - 1;
              const int kind = bid < P.mesh.n_bnd ? P.mesh.bnd_kind[bid] : MCW_BC_KILL;
              if (kind == MCW_BC_KILL)
              {
                status = MCW_ST_DEAD;
                fate = 1;
                leave = true;
              }
              else
              {
                // RT ReflectRayBC.C:68-76 + Ray::changeDirection (Ray.C:206-214)
                ElemRec<TYPE> R(P.mesh, cur);
                V3 n = mcw_side_normal<TYPE>(R, iside);
                double dn = dot(dir, n);
                V3 r = dir - (2.0 * dn) * n;
                r = r / norm(r);
                dir = unit(r);
                last = cur;
                inc_side = iside;
              }
            }
          }
        }
      }
      if (leave)
      {
        if (need_inc)
        {
          // cannot happen together with leave today (need_inc is consumed at the top), kept for safety
          ElemData<TYPE> N;
          N.load(P.mesh, cur);
          int s_in = MCW_INVALID;
#pragma unroll
          for (int k = T::NS - 1; k >= 0; --k)
            if (N.nbr[k] == last)
              s_in = k;
          inc_side = s_in;
          need_inc = false;
        }
        p.x[i] = point.x;
        p.y[i] = point.y;
        p.z[i] = point.z;
        p.ux[i] = dir.x;
        p.uy[i] = dir.y;
        p.uz[i] = dir.z;
        p.elem[i] = cur;
        p.last_elem[i] = last;
        p.inc_side[i] = (int8_t)inc_side;
        p.seed_trk[i] = s.seed_trk;
        p.seed_src[i] = s.seed_src;
        p.seed_urr[i] = s.seed_urr;
        p.sig_t[i] = s.sig_t;
        p.sig_a[i] = s.sig_a;
        p.sig_f[i] = s.sig_f;
        p.sig_nf[i] = s.sig_nf;
        p.xs_mat[i] = xs_mat;
        p.k_trk[i] = s.k_trk;
        p.n_seg[i] = n_seg;
        p.hash[i] = hash;
        p.status[i] = (int8_t)(requeue ? MCW_ST_FLY : status);
        if (fate >= 0)
          p.fate[i] = (int8_t)fate;
        if (requeue)
          q_fly_next[mcw_warp_append(cnt_next, true)] = i;
        else if (status == MCW_ST_COLLIDE)
          q_collide[mcw_warp_append(&P.counters->n_collide, true)] = i;
        else if (status == MCW_ST_FALLBACK)
          q_fallback[mcw_warp_append(&P.counters->n_fallback, true)] = i;
        active = false;
      }
    }
  }
  // segment counter: one atomic per warp
  {
    unsigned long long tot = segs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      tot += __shfl_xor_sync(MCW_FULL, tot, o);
    if (lane == 0 && tot)
      atomicAdd(&P.counters->n_segments, tot);
  }
}
