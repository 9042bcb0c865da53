// mcw_fallback.cuh -- the general element-walk step: everything TraceRay::trace does when a ray
// touches a vertex or an edge, fails the plain exit test, or sits on the domain boundary at an
// extremum.  RT TraceRay.C:564-723 (moveThroughNeighbors / moveThroughNeighbor), :725-782
// (applyOnExternalBoundary), :930-997 (findExternalBoundarySide), :1134-1246, :1387-1418;
// RT TraceRayTools.C:567-790 (atVertex / withinEdge / withinExtremaOnSide), :793-823.
// Neighbour lists come from a node -> elements CSR in ascending element id (the reference's cached
// lists depend on libMesh iteration order, RT TraceRay.C:1784-1972; differences only matter for
// exact FP ties).  This is the rare path (<1e-6 of segments for random rays): one thread walks one
// step with per-thread scratch lists; it is kept out of k_track so the hot loop stays small.
#pragma once
#include "mcw_physics.cuh"

#define MCW_MAX_NBRS 64

struct NbrInfo
{
  int elem;
  unsigned side_mask; // sides of `elem` that touch the vertex / edge / point
};

MCW_D bool mcw_fuzzy_eq(V3 a, V3 b, double tol)
{
  return fabs(a.x - b.x) + fabs(a.y - b.y) + fabs(a.z - b.z) <= tol;
}

MCW_D bool mcw_within_segment(V3 s1, V3 s2, V3 p, double tol) // RT TraceRayTools.C:793-823
{
  double len = norm(s1 - s2);
  V3 d1 = p - s1, d2 = p - s2;
  if (dot(d1, d2) > tol * len)
    return false;
  return fabs(norm(d1) + norm(d2) - len) < tol * len;
}

template <int TYPE>
MCW_D int edge_node(int e, int k);
template <>
MCW_D int edge_node<MCW_TET4>(int e, int k)
{
  // {0,1},{1,2},{0,2},{0,3},{1,3},{2,3}
  const unsigned tbl = 0u | (1u << 2) | (1u << 4) | (2u << 6) | (0u << 8) | (2u << 10) | (0u << 12) | (3u << 14) |
                       (1u << 16) | (3u << 18) | (2u << 20) | (3u << 22);
  return (tbl >> (2 * (2 * e + k))) & 3u;
}
template <>
MCW_D int edge_node<MCW_HEX8>(int e, int k)
{
  // {0,1},{1,2},{2,3},{0,3},{0,4},{1,5},{2,6},{3,7},{4,5},{5,6},{6,7},{4,7}; 3 bits x 24 = 72 bits
  const unsigned long long lo = 0ull | (1ull << 3) | (1ull << 6) | (2ull << 9) | (2ull << 12) | (3ull << 15) |
                                (0ull << 18) | (3ull << 21) | (0ull << 24) | (4ull << 27) | (1ull << 30) |
                                (5ull << 33) | (2ull << 36) | (6ull << 39) | (3ull << 42) | (7ull << 45);
  const unsigned hi = 4u | (5u << 3) | (5u << 6) | (6u << 9) | (6u << 12) | (7u << 15) | (4u << 18) | (7u << 21);
  int q = 2 * e + k;
  return q < 16 ? (int)((lo >> (3 * q)) & 7ull) : (int)((hi >> (3 * (q - 16))) & 7u);
}
template <>
MCW_D int edge_node<MCW_QUAD4>(int, int)
{
  return 0;
}

template <int TYPE>
MCW_D bool node_on_side(int v, int s)
{
#pragma unroll
  for (int i = 0; i < McwElem<TYPE>::NPS; ++i)
    if (side_node<TYPE>(s, i) == v)
      return true;
  return false;
}

struct StepResult
{
  bool failed;
  int elem;  // element the segment lies in
  int iside; // exit side of that element
  double idist;
  V3 ipoint;
};

template <int TYPE>
struct GeneralWalk
{
  using T = McwElem<TYPE>;
  int cur, last, inc_side;
  V3 point, dir;
  Extrema ext; // extrema of the previous intersection on entry to step(), of this one on exit

  MCW_D int64_t node(const McwMeshDev& m, int e, int v) const { return m.conn[(int64_t)e * T::NV + v]; }

  MCW_D int vertex_neighbors(const McwMeshDev& m, int64_t nd, NbrInfo* out) const
  {
    int n = 0;
    for (int64_t k = m.n2e_off[nd]; k < m.n2e_off[nd + 1] && n < MCW_MAX_NBRS; ++k)
    {
      int c = m.n2e[k];
      int lv = MCW_INVALID;
      for (int v = 0; v < T::NV; ++v)
        if (node(m, c, v) == nd)
          lv = v;
      unsigned mask = 0;
      for (int s = 0; s < T::NS; ++s)
        if (node_on_side<TYPE>(lv, s))
          mask |= 1u << s;
      out[n].elem = c;
      out[n++].side_mask = mask;
    }
    return n;
  }

  MCW_D int edge_neighbors(const McwMeshDev& m, int64_t n1, int64_t n2, NbrInfo* out) const
  {
    int n = 0;
    for (int64_t k = m.n2e_off[n1]; k < m.n2e_off[n1 + 1] && n < MCW_MAX_NBRS; ++k)
    {
      int c = m.n2e[k];
      int a = MCW_INVALID, b = MCW_INVALID;
      for (int v = 0; v < T::NV; ++v)
      {
        if (node(m, c, v) == n1)
          a = v;
        if (node(m, c, v) == n2)
          b = v;
      }
      if (b == MCW_INVALID)
        continue;
      unsigned mask = 0;
      for (int s = 0; s < T::NS; ++s)
        if (node_on_side<TYPE>(a, s) && node_on_side<TYPE>(b, s))
          mask |= 1u << s;
      out[n].elem = c;
      out[n++].side_mask = mask;
    }
    return n;
  }

  MCW_D int point_neighbors(const McwMeshDev& m, int e, V3 p, NbrInfo* out) const
  {
    // candidates = every element sharing a node with e, ascending id, deduplicated
    int n = 0;
    int prev = -1;
    while (n < MCW_MAX_NBRS)
    {
      int best = 0x7fffffff;
      for (int v = 0; v < T::NV; ++v)
      {
        int64_t nd = node(m, e, v);
        for (int64_t k = m.n2e_off[nd]; k < m.n2e_off[nd + 1]; ++k)
        {
          int c = m.n2e[k];
          if (c > prev && c < best)
            best = c;
        }
      }
      if (best == 0x7fffffff)
        break;
      prev = best;
      ElemRec<TYPE> R(m, best);
      if (!mcw_contains_point<TYPE>(R, p))
        continue;
      unsigned mask = 0;
      for (int s = 0; s < T::NS; ++s)
        if (mcw_side_contains_point<TYPE>(R, s, p))
          mask |= 1u << s;
      if (mask)
      {
        out[n].elem = best;
        out[n++].side_mask = mask;
      }
    }
    return n;
  }

  MCW_D int get_neighbors(const McwMeshDev& m, int e, const Extrema& x, V3 p, NbrInfo* out) const
  {
    if (!x.atExtrema())
      return point_neighbors(m, e, p, out);
    if (x.atVertex())
      return vertex_neighbors(m, node(m, e, x.first), out);
    return edge_neighbors(m, node(m, e, x.first), node(m, e, x.second), out);
  }

  MCW_D int at_vertex(const ElemRec<TYPE>& R, V3 p) const // RT TraceRayTools.C:567-575
  {
    for (int v = 0; v < T::NV; ++v)
      if (mcw_fuzzy_eq(R.P(v), p, MCW_TRACE_TOL))
        return v;
    return MCW_INVALID;
  }
  MCW_D int at_vertex_on_side(const ElemRec<TYPE>& R, V3 p, int s) const // :661-670
  {
    for (int i = 0; i < T::NPS; ++i)
      if (mcw_fuzzy_eq(R.P(side_node<TYPE>(s, i)), p, MCW_TRACE_TOL))
        return side_node<TYPE>(s, i);
    return MCW_INVALID;
  }
  MCW_D bool within_edge(const ElemRec<TYPE>& R, V3 p, Extrema& x) const // :577-597
  {
    for (int k = 0; k < T::NE; ++k)
      if (mcw_within_segment(R.P(edge_node<TYPE>(k, 0)), R.P(edge_node<TYPE>(k, 1)), p, MCW_TRACE_TOL))
      {
        x.setEdge(edge_node<TYPE>(k, 0), edge_node<TYPE>(k, 1));
        return true;
      }
    return false;
  }
  MCW_D bool within_edge_on_side(const ElemRec<TYPE>& R, V3 p, int s, Extrema& x) const // :740-765
  {
    int last_n = side_node<TYPE>(s, T::NPS - 1);
    for (int i = 0; i < T::NPS; ++i)
    {
      int sv = side_node<TYPE>(s, i);
      if (mcw_within_segment(R.P(last_n), R.P(sv), p, MCW_TRACE_TOL))
      {
        x.setEdge(last_n, sv);
        return true;
      }
      last_n = sv;
    }
    return false;
  }
  MCW_D void within_extrema_on_side(const ElemRec<TYPE>& R, V3 p, int s, Extrema& x) const // :767-790
  {
    x.first = at_vertex_on_side(R, p, s);
    x.second = MCW_INVALID;
    if (x.atVertex())
      return;
    if (T::DIM == 3)
      within_edge_on_side(R, p, s, x);
  }

  // RT TraceRay.C:677-723 moveThroughNeighbor
  MCW_D bool try_neighbor(const McwMeshDev& m, const NbrInfo& ni, int& inc, V3& p2, int& s2, Extrema& e2,
                          double& d2) const
  {
    ElemRec<TYPE> R(m, ni.elem);
    inc = MCW_INVALID;
    for (int s = 0; s < T::NS; ++s)
      if ((ni.side_mask >> s) & 1u)
        if (dot(mcw_side_normal<TYPE>(R, s), dir) < MCW_LOOSE_TRACE_TOL)
        {
          inc = s;
          break;
        }
    if (inc == MCW_INVALID)
      return false;
    e2.invalidate();
    return mcw_exits_elem<TYPE>(m, R, inc, point, dir, p2, s2, e2, d2);
  }

  // One iteration of the trace loop up to (and including) finding the next segment.
  MCW_D StepResult step(const McwMeshDev& m, unsigned long long& n_ve, unsigned long long& n_moved)
  {
    StepResult r;
    r.failed = false;
    r.iside = MCW_INVALID;
    r.idist = 0;
    r.ipoint = point;
    Extrema last_ext = ext;
    bool exits = false;
    ext.invalidate();
    if (!last_ext.atExtrema())
    {
      ElemRec<TYPE> R(m, cur);
      Extrema e2;
      if (mcw_exits_elem<TYPE>(m, R, inc_side, point, dir, r.ipoint, r.iside, e2, r.idist))
      {
        exits = true;
        ext = e2;
      }
    }
    if (!exits)
    {
      NbrInfo nb[MCW_MAX_NBRS];
      int nn = 0;
      int mtn_last;
      if (last_ext.atExtrema())
      {
        mtn_last = last;
        nn = get_neighbors(m, last, last_ext, point, nb);
      }
      else
      {
        mtn_last = cur;
        ElemRec<TYPE> R(m, cur);
        int at_v = inc_side != MCW_INVALID ? at_vertex_on_side(R, point, inc_side) : at_vertex(R, point);
        if (at_v != MCW_INVALID)
          nn = vertex_neighbors(m, node(m, cur, at_v), nb);
        else if (T::DIM == 3)
        {
          Extrema e3;
          if (inc_side != MCW_INVALID)
            within_edge_on_side(R, point, inc_side, e3);
          else
            within_edge(R, point, e3);
          if (e3.atEdge())
            nn = edge_neighbors(m, node(m, cur, e3.first), node(m, cur, e3.second), nb);
        }
        if (nn == 0)
          nn = point_neighbors(m, cur, point, nb);
        if (nn == 0)
        {
          r.failed = true;
          return r;
        }
      }
      ++n_moved;
      double longest = 1.0e-12;
      int best_elem = MCW_INVALID, best_inc = MCW_INVALID, last_idx = -1;
      for (int q = 0; q < nn; ++q)
      {
        if (nb[q].elem == mtn_last)
        {
          last_idx = q;
          continue;
        }
        int inc, s2;
        V3 p2;
        Extrema e2;
        double d2;
        if (try_neighbor(m, nb[q], inc, p2, s2, e2, d2) && d2 > longest)
        {
          best_elem = nb[q].elem;
          best_inc = inc;
          r.ipoint = p2;
          r.iside = s2;
          ext = e2;
          r.idist = d2;
          longest = d2;
        }
      }
      if (best_elem == MCW_INVALID && last_idx >= 0)
      {
        int inc, s2;
        V3 p2;
        Extrema e2;
        double d2;
        if (try_neighbor(m, nb[last_idx], inc, p2, s2, e2, d2) && d2 > longest)
        {
          best_elem = nb[last_idx].elem;
          best_inc = inc;
          r.ipoint = p2;
          r.iside = s2;
          ext = e2;
          r.idist = d2;
        }
      }
      if (best_elem == MCW_INVALID)
      {
        r.failed = true;
        return r;
      }
      last = cur;
      cur = best_elem;
      inc_side = best_inc;
    }
    if (ext.atExtrema())
      ++n_ve;
    r.elem = cur;
    return r;
  }

  // After a segment without collision: RT TraceRay.C:1387-1585.  0 = continue, 1 = killed by a
  // boundary condition, 2 = failed.
  MCW_D int cross(const McwMeshDev& m, StepResult& st)
  {
    int iside = st.iside;
    V3 ipoint = st.ipoint;
    {
      ElemRec<TYPE> R(m, cur);
      if (T::DIM > 1 && ext.atExtrema() && R.nbr(iside) >= 0)
      {
        bool on_bb = false;
        if (m.rectangular)
        {
          double tol = MCW_LOOSE_TRACE_TOL * m.domain_max_length;
          const double pc[3] = {ipoint.x, ipoint.y, ipoint.z};
          for (int d = 0; d < T::DIM; ++d)
            if (fabs(pc[d] - m.bb_min[d]) <= tol || fabs(pc[d] - m.bb_max[d]) <= tol)
              on_bb = true;
        }
        if (!m.rectangular || on_bb)
        {
          // findExternalBoundarySide
          bool found = false;
          const bool at_edge = ext.atEdge();
          for (int s = 0; s < T::NS && !found; ++s)
            if (R.nbr(s) < 0 && s != iside && node_on_side<TYPE>(ext.first, s) &&
                (!at_edge || node_on_side<TYPE>(ext.second, s)) &&
                !(dot(mcw_side_normal<TYPE>(R, s), dir) < MCW_TRACE_TOL))
            {
              iside = s;
              found = true;
            }
          if (!found)
          {
            NbrInfo nb[MCW_MAX_NBRS];
            int nn = get_neighbors(m, cur, ext, ipoint, nb);
            for (int q = 0; q < nn && !found; ++q)
            {
              if (nb[q].elem == cur)
                continue;
              ElemRec<TYPE> N(m, nb[q].elem);
              for (int s = 0; s < T::NS && !found; ++s)
                if (((nb[q].side_mask >> s) & 1u) && dot(mcw_side_normal<TYPE>(N, s), dir) > MCW_TRACE_TOL &&
                    N.nbr(s) < 0)
                {
                  Extrema be;
                  within_extrema_on_side(N, ipoint, s, be);
                  last = cur;
                  cur = nb[q].elem;
                  iside = s;
                  ext = be;
                  found = true;
                }
            }
          }
        }
      }
    }
    point = ipoint;
    ElemRec<TYPE> R(m, cur);
    const int nbid = R.nbr(iside);
    if (nbid >= 0)
    {
      ElemRec<TYPE> N(m, nbid);
      int s_in = MCW_INVALID;
      for (int k = T::NS - 1; k >= 0; --k)
        if (N.nbr(k) == cur)
          s_in = k;
      inc_side = s_in;
      last = cur;
      cur = nbid;
      return 0;
    }
    // external boundary
    int bnd_ids[8], bnd_elem[8], bnd_side[8], nb_ = 0;
    auto add = [&](int e2, int s2) {
      ElemRec<TYPE> E2(m, e2);
      int id = -E2.nbr(s2) - 1;
      for (int q = 0; q < nb_; ++q)
        if (bnd_ids[q] == id)
          return;
      if (nb_ < 8)
      {
        bnd_ids[nb_] = id;
        bnd_elem[nb_] = e2;
        bnd_side[nb_++] = s2;
      }
    };
    if (ext.atExtrema())
    {
      NbrInfo nb[MCW_MAX_NBRS];
      int nn = get_neighbors(m, cur, ext, ipoint, nb);
      for (int q = 0; q < nn; ++q)
      {
        ElemRec<TYPE> N(m, nb[q].elem);
        for (int s = 0; s < T::NS; ++s)
          if (((nb[q].side_mask >> s) & 1u) && N.nbr(s) < 0 && dot(mcw_side_normal<TYPE>(N, s), dir) > MCW_TRACE_TOL)
            add(nb[q].elem, s);
      }
    }
    else
      add(cur, iside);
    if (nb_ == 0)
      return 2;
    for (int q = 0; q < nb_; ++q)
      if (bnd_ids[q] >= m.n_bnd || m.bnd_kind[bnd_ids[q]] == MCW_BC_KILL)
        return 1;
    for (int q = 0; q < nb_; ++q)
    {
      ElemRec<TYPE> B(m, bnd_elem[q]);
      V3 n = mcw_side_normal<TYPE>(B, bnd_side[q]);
      double dn = dot(dir, n);
      V3 rr = dir - (2.0 * dn) * n;
      rr = rr / norm(rr);
      dir = unit(rr);
    }
    last = cur;
    inc_side = iside;
    return 0;
  }
};

// One general step of history i (status MCW_ST_FALLBACK), including the segment physics: the rare path of
// RT TraceRay.C:564-997,1134-1246 for exits through a vertex or an edge.  Returns the new status.
template <int TYPE>
MCW_D int mcw_fallback_history(const McwParams& P, int i)
{
  const McwParticlesDev& p = P.p;
  GeneralWalk<TYPE> W;
  W.cur = p.elem[i];
  W.last = p.last_elem[i];
  W.inc_side = p.inc_side[i];
  W.point = v3(p.x[i], p.y[i], p.z[i]);
  W.dir = v3(p.ux[i], p.uy[i], p.uz[i]);
  W.ext.first = p.ext_first[i];
  W.ext.second = p.ext_second[i];
  unsigned long long n_ve = 0, n_moved = 0;
  int status = MCW_ST_FLY, fate = -1;
  StepResult st = W.step(P.mesh, n_ve, n_moved);
  if (st.failed)
  {
    status = MCW_ST_DEAD;
    fate = 2;
    atomicAdd(&P.counters->n_failed, 1ull);
  }
  else
  {
    ElemRec<TYPE> R(P.mesh, W.cur);
    SegState s;
    s.E = p.E[i];
    s.wgt = p.wgt[i];
    s.seed_trk = p.seed_trk[i];
    s.seed_src = p.seed_src[i];
    s.seed_urr = p.seed_urr[i];
    const int slot = R.slot();
    const int mat = P.slot_mat[slot];
    {
      // every material's cross sections at this energy are in the history's sectors (mcw_lookup_all_slots)
      const Macro M = mcw_load_xs(p.xs, P.n_slots, i, slot);
      s.sig_t = M.total;
      s.sig_a = M.absorption;
      s.sig_f = M.fission;
      s.sig_nf = M.nu_fission;
    }
    s.k_trk = p.k_trk[i];
    p.n_seg[i] += 1;
    atomicAdd(&P.counters->n_segments, 1ull);
    V3 pos;
    const bool collided = mcw_segment(P, s, W.cur, R.block(), mat, st.ipoint, W.dir, st.idist, pos);
    p.hash[i] = mcw_hash_step(p.hash[i], W.cur, collided ? 15 : st.iside);
    p.seed_trk[i] = s.seed_trk;
    p.seed_src[i] = s.seed_src;
    p.seed_urr[i] = s.seed_urr;
    p.k_trk[i] = s.k_trk;
    if (collided)
    {
      W.point = pos;
      status = MCW_ST_COLLIDE;
      if (MCW_RARE && P.trace)
        p.tx[i] = st.ipoint.x, p.ty[i] = st.ipoint.y, p.tz[i] = st.ipoint.z;
    }
    else
    {
      int c = W.cross(P.mesh, st);
      if (MCW_RARE && P.trace)
        mcw_trace_emit(P, i, p.n_seg[i], 1, st.ipoint, s.E, s.wgt, p.n_progeny[i], s.seed_trk, s.seed_src, s.seed_urr);
      if (c == 1)
      {
        status = MCW_ST_DEAD;
        fate = 1;
      }
      else if (c == 2)
      {
        status = MCW_ST_DEAD;
        fate = 2;
        atomicAdd(&P.counters->n_failed, 1ull);
      }
    }
  }
  p.x[i] = W.point.x;
  p.y[i] = W.point.y;
  p.z[i] = W.point.z;
  p.ux[i] = W.dir.x;
  p.uy[i] = W.dir.y;
  p.uz[i] = W.dir.z;
  p.elem[i] = W.cur;
  p.last_elem[i] = W.last;
  p.inc_side[i] = (int8_t)W.inc_side;
  p.ext_first[i] = (int8_t)(status == MCW_ST_COLLIDE ? MCW_INVALID : W.ext.first);
  p.ext_second[i] = (int8_t)(status == MCW_ST_COLLIDE ? MCW_INVALID : W.ext.second);
  p.status[i] = (int8_t)status;
  if (fate >= 0)
    p.fate[i] = (int8_t)fate;
  if (n_ve)
    atomicAdd(&P.counters->n_vertex_edge, n_ve);
  if (n_moved)
    atomicAdd(&P.counters->n_moved, n_moved);
  return status;
}

// One general step for every particle in the fallback queue.
template <int TYPE>
__global__ void __launch_bounds__(64)
k_fallback(McwParams P, const int32_t* __restrict__ queue, const unsigned long long* __restrict__ count,
           int32_t* __restrict__ q_fly_next, unsigned long long* __restrict__ cnt_next, int32_t* __restrict__ q_collide)
{
  // grid-stride: a small fixed grid serves whatever the (device-side) count turns out to be
  const int64_t n_q = (int64_t)*count;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_q; idx += (int64_t)gridDim.x * blockDim.x)
  {
    const int i = queue[idx];
    if (P.p.status[i] != MCW_ST_FALLBACK)
      continue;
    const int status = mcw_fallback_history<TYPE>(P, i);
    if (!P.ordered_queues)
    {
      if (status == MCW_ST_FLY)
        q_fly_next[atomicAdd(cnt_next, 1ull)] = i;
      if (status == MCW_ST_COLLIDE)
        q_collide[atomicAdd(&P.counters->n_collide, 1ull)] = i;
    }
  }
}
