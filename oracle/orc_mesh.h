// orc_mesh.h -- CPU ORACLE (test infrastructure): mesh + element walk.
//
// Restates the MOOSE ray-tracing element walk for QUAD4 / TET4 / HEX8:
//   RT src/raytracing/TraceRay.C:292-493 (exitsElem), :564-723 (moveThroughNeighbors),
//   :725-782 (applyOnExternalBoundary), :930-997 (findExternalBoundarySide), :999-1592 (trace);
//   RT src/raytracing/TraceRayTools.C:60-128 (lineLineIntersect2D), :371-553 (intersectTriangle /
//   intersectQuad), :567-790 (atVertex / withinEdge), :793-823 (isWithinSegment);
//   RT include/raytracing/TraceRayTools.h:558-856 (sideIntersectedByLine per element type);
//   RT src/raybcs/ReflectRayBC.C:39-77; RT src/raytracing/ClaimRays.C:104-163.
// libMesh itself is not in the reference tree; the element conventions (side/edge node maps, hmax,
// centroid face normals, contains_point tolerance) are restated from SURVEY.md Appendix B.
// Where the reference's answer depends on libMesh-internal iteration order (the order of cached
// vertex/edge neighbour lists, RT TraceRay.C:1784-1972) this restatement uses ascending element id;
// such cases only matter for exact FP ties and are counted (n_vertex_edge_hits).
#pragma once
#include "orc_core.h"

namespace orc
{
struct ElemTopo
{
  int nv, ns, nps, ne; // vertices, sides, nodes per side, edges
  int side_nodes[6][4];
  int edge_nodes[12][2];
};

inline const ElemTopo& topo(int type)
{
  static const ElemTopo Q4 = {4, 4, 2, 0, {{0, 1}, {1, 2}, {2, 3}, {3, 0}}, {}};
  static const ElemTopo T4 = {4,
                              4,
                              3,
                              6,
                              {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {2, 0, 3}},
                              {{0, 1}, {1, 2}, {0, 2}, {0, 3}, {1, 3}, {2, 3}}};
  static const ElemTopo H8 = {
      8,
      6,
      4,
      12,
      {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}},
      {{0, 1}, {1, 2}, {2, 3}, {0, 3}, {0, 4}, {1, 5}, {2, 6}, {3, 7}, {4, 5}, {5, 6}, {6, 7}, {4, 7}}};
  return type == MCW_QUAD4 ? Q4 : (type == MCW_TET4 ? T4 : H8);
}

constexpr int INVALID = -1;

struct Extrema // RT include/raytracing/ElemExtrema.h
{
  int first = INVALID, second = INVALID;
  bool atVertex() const { return first != INVALID && second == INVALID; }
  bool atEdge() const { return first != INVALID && second != INVALID; }
  bool atExtrema() const { return first != INVALID; }
  void invalidate() { first = second = INVALID; }
  void setVertex(int v)
  {
    first = v;
    second = INVALID;
  }
  void setEdge(int a, int b)
  {
    first = a;
    second = b;
  }
  bool atEdgeOf(int a, int b) const
  {
    return atEdge() && ((first == a && second == b) || (first == b && second == a));
  }
};

struct NeighborInfo // RT include/raytracing/NeighborInfo.h
{
  int elem;
  int n_sides;
  int sides[6];
  V3 normals[6];
};

struct Mesh
{
  int type = MCW_HEX8, dim = 3;
  int64_t n_nodes = 0, n_elem = 0;
  std::vector<V3> xyz;
  std::vector<int64_t> conn;  // n_elem x nv
  std::vector<int32_t> block; // block index (0..n_blocks-1) per element
  std::vector<int64_t> nbr;   // n_elem x ns, -1 on boundary
  std::vector<int32_t> bnd;   // n_elem x ns boundary id, -1 interior
  std::vector<int32_t> bnd_kind;
  std::vector<double> block_hmax;
  std::vector<int64_t> n2e_off, n2e; // node -> elements CSR (ascending element id)
  V3 bb_min, bb_max;
  double domain_max_length = 0;
  bool rectangular = true;

  const ElemTopo& T() const { return topo(type); }
  V3 P(int64_t e, int v) const { return xyz[conn[e * T().nv + v]]; }
  int64_t node(int64_t e, int v) const { return conn[e * T().nv + v]; }

  void finalize(int n_blocks)
  {
    const ElemTopo& t = T();
    dim = (type == MCW_QUAD4) ? 2 : 3;
    // node -> elements
    n2e_off.assign(n_nodes + 1, 0);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int v = 0; v < t.nv; ++v)
        n2e_off[node(e, v) + 1]++;
    for (int64_t i = 0; i < n_nodes; ++i)
      n2e_off[i + 1] += n2e_off[i];
    n2e.assign(n2e_off[n_nodes], 0);
    std::vector<int64_t> fill(n2e_off.begin(), n2e_off.end() - 1);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int v = 0; v < t.nv; ++v)
        n2e[fill[node(e, v)]++] = e;
    // subdomain hmax: RT RayTracingStudy.C:497-513 (max elem->hmax() per block)
    block_hmax.assign(n_blocks, 0.0);
    for (int64_t e = 0; e < n_elem; ++e)
    {
      double h2 = 0;
      for (int a = 0; a < t.nv; ++a)
        for (int b = a + 1; b < t.nv; ++b)
        {
          V3 d = P(e, a) - P(e, b);
          h2 = std::max(h2, d.x * d.x + d.y * d.y + d.z * d.z);
        }
      block_hmax[block[e]] = std::max(block_hmax[block[e]], std::sqrt(h2));
    }
    bb_min = {INFINITY, INFINITY, INFINITY};
    bb_max = {-INFINITY, -INFINITY, -INFINITY};
    for (auto& p : xyz)
    {
      bb_min = {std::min(bb_min.x, p.x), std::min(bb_min.y, p.y), std::min(bb_min.z, p.z)};
      bb_max = {std::max(bb_max.x, p.x), std::max(bb_max.y, p.y), std::max(bb_max.z, p.z)};
    }
    domain_max_length = 1.01 * norm(bb_max - bb_min); // RT RayTracingStudy.C:154
    // rectangular domain: every boundary-side vertex lies on the bounding box
    rectangular = true;
    double tol = 1e-10 * domain_max_length;
    for (int64_t e = 0; e < n_elem && rectangular; ++e)
      for (int s = 0; s < t.ns && rectangular; ++s)
        if (nbr[e * t.ns + s] < 0)
        {
          bool on_plane = false;
          for (int d = 0; d < dim && !on_plane; ++d)
          {
            bool lo = true, hi = true;
            for (int i = 0; i < t.nps; ++i)
            {
              V3 p = P(e, t.side_nodes[s][i]);
              double c = d == 0 ? p.x : (d == 1 ? p.y : p.z);
              double mn = d == 0 ? bb_min.x : (d == 1 ? bb_min.y : bb_min.z);
              double mx = d == 0 ? bb_max.x : (d == 1 ? bb_max.y : bb_max.z);
              lo = lo && std::fabs(c - mn) <= tol;
              hi = hi && std::fabs(c - mx) <= tol;
            }
            on_plane = lo || hi;
          }
          if (!on_plane)
            rectangular = false;
        }
  }

  bool nodeOnSide(int v, int s) const
  {
    const ElemTopo& t = T();
    for (int i = 0; i < t.nps; ++i)
      if (t.side_nodes[s][i] == v)
        return true;
    return false;
  }

  // Outward unit normal at the side centroid (libMesh FE face normal, RT RayTracingStudy.C:1551-1572)
  V3 sideNormal(int64_t e, int s) const
  {
    const ElemTopo& t = T();
    if (type == MCW_QUAD4)
    {
      V3 a = P(e, t.side_nodes[s][0]), b = P(e, t.side_nodes[s][1]);
      V3 n = {b.y - a.y, -(b.x - a.x), 0.0};
      return unit(n);
    }
    if (type == MCW_TET4)
    {
      V3 a = P(e, t.side_nodes[s][0]), b = P(e, t.side_nodes[s][1]), c = P(e, t.side_nodes[s][2]);
      return unit(cross(b - a, c - a));
    }
    V3 a = P(e, t.side_nodes[s][0]), b = P(e, t.side_nodes[s][1]), c = P(e, t.side_nodes[s][2]),
       d = P(e, t.side_nodes[s][3]);
    V3 dxi = (b + c) - (a + d), deta = (c + d) - (a + b);
    return unit(cross(dxi, deta));
  }

  V3 sideCentroid(int64_t e, int s) const
  {
    const ElemTopo& t = T();
    V3 c = {0, 0, 0};
    for (int i = 0; i < t.nps; ++i)
      c = c + P(e, t.side_nodes[s][i]);
    return c / (double)t.nps;
  }

  double elemHmax(int64_t e) const { return block_hmax[block[e]]; }

  // Elem::contains_point restated for convex planar-sided elements: inside every side plane within
  // libMesh TOLERANCE (1e-6) scaled by the block's hmax.
  bool containsPoint(int64_t e, V3 p) const
  {
    const ElemTopo& t = T();
    double tol = 1e-6 * elemHmax(e);
    for (int s = 0; s < t.ns; ++s)
    {
      V3 n = sideNormal(e, s);
      V3 d = p - sideCentroid(e, s);
      if (dim == 2)
        d.z = 0;
      if (dot(d, n) > tol)
        return false;
    }
    return true;
  }
  bool sideContainsPoint(int64_t e, int s, V3 p) const
  {
    double tol = 1e-6 * elemHmax(e);
    V3 d = p - sideCentroid(e, s);
    if (dim == 2)
      d.z = 0;
    return std::fabs(dot(d, sideNormal(e, s))) <= tol && containsPoint(e, p);
  }
};

// ---------------------------------------------------------------- geometric primitives
inline bool fuzzyEq(V3 a, V3 b, double tol) // libMesh TypeVector::absolute_fuzzy_equals (L1 norm)
{
  return std::fabs(a.x - b.x) + std::fabs(a.y - b.y) + std::fabs(a.z - b.z) <= tol;
}

// RT TraceRayTools.C:813-823
inline bool isWithinSegment(V3 s1, V3 s2, double len, V3 p, double tol = TRACE_TOLERANCE)
{
  V3 d1 = p - s1, d2 = p - s2;
  if (dot(d1, d2) > tol * len)
    return false;
  return std::fabs(norm(d1) + norm(d2) - len) < tol * len;
}
inline bool isWithinSegment(V3 s1, V3 s2, V3 p, double tol = TRACE_TOLERANCE)
{
  return isWithinSegment(s1, s2, norm(s1 - s2), p, tol);
}

// RT TraceRayTools.C:371-478
inline bool intersectTriangle(V3 start, V3 dir, const Mesh& m, int64_t e, int v0, int v1, int v2,
                              double& dist, Extrema& ext, double hmax)
{
  const double inv_hmax = 1.0 / hmax;
  V3 p0 = m.P(e, v0);
  V3 edge1 = (m.P(e, v1) - p0) * inv_hmax;
  V3 edge2 = (m.P(e, v2) - p0) * inv_hmax;
  V3 pvec = cross(dir, edge2);
  double det = dot(edge1, pvec);
  if (det < TRACE_TOLERANCE)
    return false;
  V3 tvec = (start - p0) * inv_hmax;
  double u = dot(tvec, pvec);
  if (u < -TRACE_TOLERANCE || u > det + TRACE_TOLERANCE)
    return false;
  V3 qvec = cross(tvec, edge1);
  double v = dot(dir, qvec);
  if (v < -TRACE_TOLERANCE || u + v > det + TRACE_TOLERANCE)
    return false;
  double possible = dot(edge2, qvec) / det;
  if (possible <= TRACE_TOLERANCE)
    return false;
  dist = possible * hmax;
  if (u < TRACE_TOLERANCE * det)
  {
    if (v < TRACE_TOLERANCE * det)
      ext.setVertex(v0);
    else if (v > (1.0 - TRACE_TOLERANCE) * det)
      ext.setVertex(v2);
    else
      ext.setEdge(v0, v2);
  }
  else if (v < TRACE_TOLERANCE * det)
  {
    if (u > (1.0 - TRACE_TOLERANCE) * det)
      ext.setVertex(v1);
    else
      ext.setEdge(v0, v1);
  }
  else if (u + v > (1.0 - TRACE_TOLERANCE) * det)
    ext.setEdge(v1, v2);
  return true;
}

// RT TraceRayTools.C:480-553
inline bool intersectQuad(V3 start, V3 dir, const Mesh& m, int64_t e, int v00, int v10, int v11,
                          int v01, double& dist, Extrema& ext, double hmax)
{
  bool hit = intersectTriangle(start, dir, m, e, v00, v10, v11, dist, ext, hmax);
  if (!hit)
    hit = intersectTriangle(start, dir, m, e, v11, v01, v00, dist, ext, hmax);
  if (hit && ext.atEdgeOf(v00, v11))
    ext.invalidate();
  return hit;
}

// RT TraceRayTools.C:60-128; returns segment vertex 0/1 or -1 in seg_v
inline bool lineLineIntersect2D(V3 start, V3 dir, double length, V3 v0, V3 v1, V3& ipoint,
                                double& idist, int& seg_v)
{
  V3 r = dir * length;
  V3 s = v1 - v0;
  double rxs = r.x * s.y - r.y * s.x;
  if (std::fabs(rxs) < TRACE_TOLERANCE)
    return false;
  V3 v0mu0 = v0 - start;
  double t = (v0mu0.x * s.y - v0mu0.y * s.x) / rxs;
  if (0 >= t + TRACE_TOLERANCE || t - TRACE_TOLERANCE > 1.0)
    return false;
  double u = (v0mu0.x * r.y - v0mu0.y * r.x) / rxs;
  if (0 < u + TRACE_TOLERANCE && u - TRACE_TOLERANCE <= 1.0)
  {
    ipoint = start + r * t;
    idist = t * length;
    if (u < TRACE_TOLERANCE)
      seg_v = 0;
    else if (u > 1.0 - TRACE_TOLERANCE)
      seg_v = 1;
    return true;
  }
  return false;
}

// RT TraceRayTools.h:558-856 sideIntersectedByLine<T>
inline bool sideIntersectedByLine(const Mesh& m, int64_t e, V3 start, V3 dir, int s, V3& ipoint,
                                  double& idist, Extrema& ext, double hmax)
{
  const ElemTopo& t = m.T();
  if (m.type == MCW_QUAD4)
  {
    int seg_v = -1;
    bool hit = lineLineIntersect2D(start, dir, m.domain_max_length, m.P(e, t.side_nodes[s][0]),
                                   m.P(e, t.side_nodes[s][1]), ipoint, idist, seg_v);
    if (seg_v >= 0)
      ext.setVertex(t.side_nodes[s][seg_v]);
    return hit;
  }
  bool hit;
  if (m.type == MCW_TET4)
    hit = intersectTriangle(start, dir, m, e, t.side_nodes[s][2], t.side_nodes[s][1],
                            t.side_nodes[s][0], idist, ext, hmax);
  else
    hit = intersectQuad(start, dir, m, e, t.side_nodes[s][3], t.side_nodes[s][2],
                        t.side_nodes[s][1], t.side_nodes[s][0], idist, ext, hmax);
  if (hit)
    ipoint = start + idist * dir;
  return hit;
}

// RT TraceRay.C:292-493 without backface culling and with planar sides (MaCaw inputs use neither)
inline bool exitsElem(const Mesh& m, int64_t e, int incoming_side, V3 start, V3 dir, V3& ipoint,
                      int& iside, Extrema& iext, double& idist, int64_t* n_calls = nullptr)
{
  const ElemTopo& t = m.T();
  const double hmax = m.elemHmax(e);
  double best = TRACE_TOLERANCE * hmax;
  iside = INVALID;
  if (n_calls)
    ++*n_calls;
  for (int s = 0; s < t.ns; ++s)
  {
    if (s == incoming_side)
      continue;
    V3 cp = {0, 0, 0};
    double cd = 0;
    Extrema ce;
    if (sideIntersectedByLine(m, e, start, dir, s, cp, cd, ce, hmax) && cd > best)
    {
      iside = s;
      idist = cd;
      ipoint = cp;
      iext = ce;
      best = cd;
    }
  }
  return iside != INVALID;
}

// RT TraceRayTools.C:567-575, :661-670
inline int atVertex(const Mesh& m, int64_t e, V3 p)
{
  for (int v = 0; v < m.T().nv; ++v)
    if (fuzzyEq(m.P(e, v), p, TRACE_TOLERANCE))
      return v;
  return INVALID;
}
inline int atVertexOnSide(const Mesh& m, int64_t e, V3 p, int s)
{
  const ElemTopo& t = m.T();
  for (int i = 0; i < t.nps; ++i)
    if (fuzzyEq(m.P(e, t.side_nodes[s][i]), p, TRACE_TOLERANCE))
      return t.side_nodes[s][i];
  return INVALID;
}
// RT TraceRayTools.C:577-597, :740-765
inline bool withinEdge(const Mesh& m, int64_t e, V3 p, Extrema& ext)
{
  const ElemTopo& t = m.T();
  for (int k = 0; k < t.ne; ++k)
    if (isWithinSegment(m.P(e, t.edge_nodes[k][0]), m.P(e, t.edge_nodes[k][1]), p))
    {
      ext.setEdge(t.edge_nodes[k][0], t.edge_nodes[k][1]);
      return true;
    }
  return false;
}
inline bool withinEdgeOnSide(const Mesh& m, int64_t e, V3 p, int s, Extrema& ext)
{
  const ElemTopo& t = m.T();
  int last_n = t.side_nodes[s][t.nps - 1];
  for (int i = 0; i < t.nps; ++i)
  {
    if (isWithinSegment(m.P(e, last_n), m.P(e, t.side_nodes[s][i]), p))
    {
      ext.setEdge(last_n, t.side_nodes[s][i]);
      return true;
    }
    last_n = t.side_nodes[s][i];
  }
  return false;
}
inline bool withinExtremaOnSide(const Mesh& m, int64_t e, V3 p, int s, Extrema& ext)
{
  ext.first = atVertexOnSide(m, e, p, s);
  ext.second = INVALID;
  if (ext.atVertex())
    return true;
  if (m.dim == 3 && withinEdgeOnSide(m, e, p, s, ext))
    return true;
  return false;
}

// Neighbour lists (RT TraceRay.C:1784-1972 + TraceRayTools.C:130-331), canonical ascending-id order.
inline void vertexNeighbors(const Mesh& m, int64_t node, std::vector<NeighborInfo>& out)
{
  const ElemTopo& t = m.T();
  out.clear();
  for (int64_t k = m.n2e_off[node]; k < m.n2e_off[node + 1]; ++k)
  {
    int64_t c = m.n2e[k];
    int lv = INVALID;
    for (int v = 0; v < t.nv; ++v)
      if (m.node(c, v) == node)
        lv = v;
    NeighborInfo ni;
    ni.elem = (int)c;
    ni.n_sides = 0;
    for (int s = 0; s < t.ns; ++s)
      if (m.nodeOnSide(lv, s))
      {
        ni.sides[ni.n_sides] = s;
        ni.normals[ni.n_sides++] = m.sideNormal(c, s);
      }
    out.push_back(ni);
  }
}
inline void edgeNeighbors(const Mesh& m, int64_t n1, int64_t n2, std::vector<NeighborInfo>& out)
{
  const ElemTopo& t = m.T();
  out.clear();
  for (int64_t k = m.n2e_off[n1]; k < m.n2e_off[n1 + 1]; ++k)
  {
    int64_t c = m.n2e[k];
    int a = INVALID, b = INVALID;
    for (int v = 0; v < t.nv; ++v)
    {
      if (m.node(c, v) == n1)
        a = v;
      if (m.node(c, v) == n2)
        b = v;
    }
    if (b == INVALID)
      continue;
    NeighborInfo ni;
    ni.elem = (int)c;
    ni.n_sides = 0;
    for (int s = 0; s < t.ns; ++s)
      if (m.nodeOnSide(a, s) && m.nodeOnSide(b, s))
      {
        ni.sides[ni.n_sides] = s;
        ni.normals[ni.n_sides++] = m.sideNormal(c, s);
      }
    out.push_back(ni);
  }
}
inline void pointNeighbors(const Mesh& m, int64_t e, V3 p, std::vector<NeighborInfo>& out)
{
  const ElemTopo& t = m.T();
  out.clear();
  std::vector<int64_t> cand;
  for (int v = 0; v < t.nv; ++v)
  {
    int64_t nd = m.node(e, v);
    for (int64_t k = m.n2e_off[nd]; k < m.n2e_off[nd + 1]; ++k)
      cand.push_back(m.n2e[k]);
  }
  std::sort(cand.begin(), cand.end());
  cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
  for (int64_t c : cand)
  {
    if (!m.containsPoint(c, p))
      continue;
    NeighborInfo ni;
    ni.elem = (int)c;
    ni.n_sides = 0;
    for (int s = 0; s < t.ns; ++s)
      if (m.sideContainsPoint(c, s, p))
      {
        ni.sides[ni.n_sides] = s;
        ni.normals[ni.n_sides++] = m.sideNormal(c, s);
      }
    if (ni.n_sides)
      out.push_back(ni);
  }
}
inline void getNeighbors(const Mesh& m, int64_t e, const Extrema& ext, V3 p,
                         std::vector<NeighborInfo>& out)
{
  if (!ext.atExtrema())
    pointNeighbors(m, e, p, out);
  else if (ext.atVertex())
    vertexNeighbors(m, m.node(e, ext.first), out);
  else
    edgeNeighbors(m, m.node(e, ext.first), m.node(e, ext.second), out);
}

// RT ClaimRays.C:104-139 claimPoint + :141-163 postClaimRay.  `hint` is any element containing the
// point (the point locator's answer); candidates are its point neighbours.
inline int64_t claimPoint(const Mesh& m, V3 p, int64_t ray_id, int64_t hint)
{
  const ElemTopo& t = m.T();
  const bool smallest = (ray_id % 2) == 0;
  int64_t best = hint;
  for (int v = 0; v < t.nv; ++v)
  {
    int64_t nd = m.node(hint, v);
    for (int64_t k = m.n2e_off[nd]; k < m.n2e_off[nd + 1]; ++k)
    {
      int64_t c = m.n2e[k];
      if ((smallest && c < best) || (!smallest && c > best))
        if (m.containsPoint(c, p))
          best = c;
    }
  }
  return best;
}
inline int startingIncomingSide(const Mesh& m, int64_t e, V3 p, V3 dir)
{
  for (int s = 0; s < m.T().ns; ++s)
    if (m.sideContainsPoint(e, s, p) && dot(m.sideNormal(e, s), dir) < TRACE_TOLERANCE)
      return s;
  return INVALID;
}

// ---------------------------------------------------------------- the trace loop
// Segment callback: returns 0 = continue, 1 = killed, 2 = trajectory changed (new start/direction
// written to *start / *dir).  Mirrors RayKernelBase::onSegment (RT RayKernelBase.h:55) as driven by
// TraceRay::onSegment (RT TraceRay.C:1987-2041).
struct SegmentCtx
{
  int64_t elem;
  int side; // exit side or INVALID
  V3 start, end;
  double length;
  V3 dir; // current ray direction (changes on reflection)
};

struct TraceStats
{
  int64_t segments = 0, vertex_edge_hits = 0, moved_through_neighbors = 0, failed = 0,
          intersection_calls = 0, reflections = 0;
};

inline uint64_t hashStep(uint64_t h, int64_t elem, int side)
{
  uint64_t key = ((uint64_t)elem << 4) | (uint64_t)(side & 15);
  return (h ^ key) * 0x100000001b3ull;
}

enum TraceEnd
{
  END_KERNEL_KILL = 0,
  END_BC_KILL = 1,
  END_FAILED = 2
};

template <typename OnSegment>
inline int trace(const Mesh& m, int64_t elem, int incoming_side, V3 point, V3 dir, OnSegment&& seg,
                 TraceStats& st, uint64_t* hash, double* distance, int64_t max_segments,
                 int64_t* last_elem_out = nullptr, double max_distance = INFINITY)
{
  double travelled = 0;
  const ElemTopo& t = m.T();
  int64_t cur = elem, last = INVALID;
  int inc_side = incoming_side;
  V3 inc_point = point;
  Extrema ext;
  std::vector<NeighborInfo> nbrs;
  for (int64_t it = 0; it < max_segments; ++it)
  {
    Extrema last_ext = ext;
    bool exits = false;
    V3 ipoint = {0, 0, 0};
    int iside = INVALID;
    double idist = 0;
    ext.invalidate();
    if (!last_ext.atExtrema())
    {
      Extrema e2;
      if (exitsElem(m, cur, inc_side, inc_point, dir, ipoint, iside, e2, idist,
                    &st.intersection_calls))
      {
        exits = true;
        ext = e2;
      }
    }
    if (!exits)
    {
      // RT TraceRay.C:1134-1246
      int64_t mtn_last;
      if (last_ext.atExtrema())
      {
        mtn_last = last;
        getNeighbors(m, last, last_ext, inc_point, nbrs);
      }
      else
      {
        mtn_last = cur;
        nbrs.clear();
        int at_v = inc_side != INVALID ? atVertexOnSide(m, cur, inc_point, inc_side)
                                       : atVertex(m, cur, inc_point);
        if (at_v != INVALID)
          vertexNeighbors(m, m.node(cur, at_v), nbrs);
        else if (m.dim == 3)
        {
          Extrema e3;
          if (inc_side != INVALID)
            withinEdgeOnSide(m, cur, inc_point, inc_side, e3);
          else
            withinEdge(m, cur, inc_point, e3);
          if (e3.atEdge())
            edgeNeighbors(m, m.node(cur, e3.first), m.node(cur, e3.second), nbrs);
        }
        if (nbrs.empty())
          pointNeighbors(m, cur, inc_point, nbrs);
        if (nbrs.empty())
        {
          ++st.failed;
          if (last_elem_out)
            *last_elem_out = cur;
          return END_FAILED;
        }
      }
      // RT TraceRay.C:564-675 moveThroughNeighbors
      ++st.moved_through_neighbors;
      double longest = 1.0e-12;
      int64_t best_elem = INVALID;
      int best_inc = INVALID;
      const NeighborInfo* last_info = nullptr;
      auto try_neighbor = [&](const NeighborInfo& ni, int& inc, V3& p2, int& s2, Extrema& e2,
                              double& d2) -> bool {
        inc = INVALID;
        for (int i = 0; i < ni.n_sides; ++i)
          if (dot(ni.normals[i], dir) < LOOSE_TRACE_TOLERANCE)
          {
            inc = ni.sides[i];
            break;
          }
        if (inc == INVALID)
          return false;
        e2.invalidate();
        return exitsElem(m, ni.elem, inc, inc_point, dir, p2, s2, e2, d2, &st.intersection_calls);
      };
      for (const NeighborInfo& ni : nbrs)
      {
        if (ni.elem == mtn_last)
        {
          last_info = &ni;
          continue;
        }
        int inc, s2;
        V3 p2;
        Extrema e2;
        double d2;
        if (try_neighbor(ni, inc, p2, s2, e2, d2) && d2 > longest)
        {
          best_elem = ni.elem;
          best_inc = inc;
          ipoint = p2;
          iside = s2;
          ext = e2;
          idist = d2;
          longest = d2;
        }
      }
      if (best_elem == INVALID && last_info)
      {
        int inc, s2;
        V3 p2;
        Extrema e2;
        double d2;
        if (try_neighbor(*last_info, inc, p2, s2, e2, d2) && d2 > longest)
        {
          best_elem = last_info->elem;
          best_inc = inc;
          ipoint = p2;
          iside = s2;
          ext = e2;
          idist = d2;
        }
      }
      if (best_elem == INVALID)
      {
        ++st.failed;
        if (last_elem_out)
          *last_elem_out = cur;
        return END_FAILED;
      }
      last = cur;
      cur = best_elem;
      inc_side = best_inc;
    }
    if (ext.atExtrema())
      ++st.vertex_edge_hits;
    ++st.segments;
    // RT TraceRay.C:1268-1315: stop at / move back to the ray's maximum distance
    bool at_end = false;
    if (max_distance == INFINITY)
      at_end = false;
    else if (std::fabs(travelled + idist - max_distance) <= 1e-10 * std::fmax(1.0, max_distance))
      at_end = true;
    else if (travelled + idist > max_distance)
    {
      double diff = travelled + idist - max_distance;
      ipoint = ipoint - dir * diff;
      idist -= diff;
      iside = INVALID;
      ext.invalidate();
      at_end = true;
    }
    travelled += idist;
    if (distance)
      *distance += idist;

    // RT TraceRay.C:1318-1364 onSegment + trajectory change
    SegmentCtx ctx{cur, iside, inc_point, ipoint, idist, dir};
    V3 nstart = inc_point, ndir = dir;
    int action = seg(ctx, nstart, ndir);
    if (hash)
      *hash = hashStep(*hash, cur, action == 0 ? iside : 15);
    if (action == 1 || at_end)
    {
      if (last_elem_out)
        *last_elem_out = cur;
      return END_KERNEL_KILL;
    }
    if (action == 2)
    {
      travelled -= idist;
      if (distance)
        *distance -= idist; // RT TraceRay.C:1356-1358 (MOOSE bug F7: segment length zeroed)
      inc_point = nstart;
      dir = ndir;
      inc_side = INVALID;
      ext.invalidate();
      continue;
    }

    // RT TraceRay.C:1387-1418: vertex/edge hit that is really on the domain boundary
    if (m.dim > 1 && ext.atExtrema() && m.nbr[cur * t.ns + iside] >= 0)
    {
      bool on_bb = false;
      if (m.rectangular)
      {
        double tol = LOOSE_TRACE_TOLERANCE * m.domain_max_length;
        const double pc[3] = {ipoint.x, ipoint.y, ipoint.z};
        const double lo[3] = {m.bb_min.x, m.bb_min.y, m.bb_min.z};
        const double hi[3] = {m.bb_max.x, m.bb_max.y, m.bb_max.z};
        for (int d = 0; d < m.dim; ++d)
          if (std::fabs(pc[d] - lo[d]) <= tol || std::fabs(pc[d] - hi[d]) <= tol)
            on_bb = true;
      }
      if (!m.rectangular || on_bb)
      {
        // RT TraceRay.C:930-997 findExternalBoundarySide
        bool found = false;
        const bool at_edge = ext.atEdge();
        for (int s = 0; s < t.ns && !found; ++s)
          if (m.nbr[cur * t.ns + s] < 0 && s != iside && m.nodeOnSide(ext.first, s) &&
              (!at_edge || m.nodeOnSide(ext.second, s)) &&
              !(dot(m.sideNormal(cur, s), dir) < TRACE_TOLERANCE))
          {
            iside = s;
            found = true;
          }
        if (!found)
        {
          getNeighbors(m, cur, ext, ipoint, nbrs);
          for (const NeighborInfo& ni : nbrs)
          {
            if (ni.elem == cur)
              continue;
            for (int i = 0; i < ni.n_sides && !found; ++i)
              if (dot(ni.normals[i], dir) > TRACE_TOLERANCE &&
                  m.nbr[(int64_t)ni.elem * t.ns + ni.sides[i]] < 0)
              {
                Extrema be;
                withinExtremaOnSide(m, ni.elem, ipoint, ni.sides[i], be);
                last = cur;
                cur = ni.elem;
                iside = ni.sides[i];
                ext = be;
                found = true;
              }
            if (found)
              break;
          }
        }
      }
    }

    inc_point = ipoint;
    int64_t nb = m.nbr[cur * t.ns + iside];
    if (nb >= 0)
    {
      int s_in = INVALID;
      for (int s = 0; s < t.ns; ++s)
        if (m.nbr[nb * t.ns + s] == cur)
        {
          s_in = s;
          break;
        }
      inc_side = s_in;
      last = cur;
      cur = nb;
    }
    else
    {
      // RT TraceRay.C:725-782 applyOnExternalBoundary + :2043-2146 onBoundary
      int bnd_ids[8], bnd_elem[8], bnd_side[8], nb_ = 0;
      auto add = [&](int64_t e2, int s2) {
        int id = m.bnd[e2 * t.ns + s2];
        for (int i = 0; i < nb_; ++i)
          if (bnd_ids[i] == id)
            return;
        if (nb_ < 8)
        {
          bnd_ids[nb_] = id;
          bnd_elem[nb_] = (int)e2;
          bnd_side[nb_++] = s2;
        }
      };
      if (m.dim != 1 && ext.atExtrema())
      {
        getNeighbors(m, cur, ext, ipoint, nbrs);
        for (const NeighborInfo& ni : nbrs)
          for (int i = 0; i < ni.n_sides; ++i)
            if (m.nbr[(int64_t)ni.elem * t.ns + ni.sides[i]] < 0 &&
                dot(ni.normals[i], dir) > TRACE_TOLERANCE)
              add(ni.elem, ni.sides[i]);
      }
      else
        add(cur, iside);
      bool kill = false;
      for (int i = 0; i < nb_; ++i)
        if (bnd_ids[i] < 0 || m.bnd_kind[bnd_ids[i]] == MCW_BC_KILL)
          kill = true;
      if (kill || nb_ == 0)
      {
        if (last_elem_out)
          *last_elem_out = cur;
        if (nb_ == 0)
        {
          ++st.failed;
          return END_FAILED;
        }
        return END_BC_KILL;
      }
      for (int i = 0; i < nb_; ++i)
      {
        // RT ReflectRayBC.C:68-76 + Ray::changeDirection (Ray.C:206-214)
        V3 n = m.sideNormal(bnd_elem[i], bnd_side[i]);
        V3 r = dir;
        double dn = dot(r, n);
        r = r - (2.0 * dn) * n;
        r = r / norm(r);
        dir = unit(r);
        ++st.reflections;
      }
      last = cur;
      inc_side = iside;
    }
  }
  ++st.failed;
  if (last_elem_out)
    *last_elem_out = cur;
  return END_FAILED;
}

} // namespace orc
