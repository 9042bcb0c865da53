// mcw_pump.cuh -- the persistent generation kernel: event-based transport without kernel boundaries.
//
// Event rounds (k_track -> k_collide -> refresh) are the efficient shape while hundreds of thousands of
// histories fly: every kernel runs convergent code over a full queue.  Below that a round lasts as long
// as its longest flight and its slowest kernel's latency floor, ~10 launches each, and the last few
// thousand histories need > 200 more collisions.  k_pump keeps the event-based split -- tracking lanes
// never execute collision code, collisions are processed 32 at a time -- but connects the two by
// device-side queues inside ONE launch:
//
//   fly ring      <- survivors of collisions, served general steps     -> tracking warps (lane refill)
//   collide ring  <- histories stopped at a collision point            -> collision warps (32 per batch)
//   fallback ring <- exits through a vertex / edge (RT TraceRay.C:564) -> a few general-walk warps
//
// A ring entry is a history index; -1 = empty.  Producers reserve a ticket with one warp-aggregated
// atomicAdd on the ring's tail, release their state writes (__threadfence) and store the index.
// Consumers reserve tickets on the ring's head for their idle lanes and POLL their own slot without
// blocking: a tracking warp keeps walking its other lanes while some wait.  Tickets and entries are
// issued in the same order, so every entry has exactly one taker.  `alive` counts histories not yet
// dead; every warp leaves when it reaches zero.  No host round trip, no grid-wide barrier, no queue
// compaction: a history advances at the pace of its own dependent chain.
//
// Per-history arithmetic is the event kernels' (same device functions), so results stay bit-identical;
// only the order of tally / bank atomics differs, as it does between any two launches.
#pragma once
#include "mcw_track.cuh"

#ifndef MCW_PUMP_MIN_BLOCKS
#define MCW_PUMP_MIN_BLOCKS 4 // 128 registers: 16 warps per SM
#endif
#define MCW_PUMP_FB_BLOCKS 16 // CTAs whose last warp serves the general-walk ring
#define MCW_PUMP_COOP_MAX 12  // refresh: fewer survivors than this in a batch -> the warp shares each history's lookups

#define MCW_PUMP_IDLE_LIMIT 4000000 // ~1 s of a warp polling empty rings: a protocol bug must not hang the device

MCW_D long long mcw_ld_alive(const long long* alive)
{
  return *reinterpret_cast<const volatile long long*>(alive);
}

// A warp that has seen nothing to do for MCW_PUMP_IDLE_LIMIT polls while histories are still counted alive
// gives up for everybody: `alive` goes far below zero, every warp leaves, the host reports the failure.
MCW_D void mcw_pump_idle(long long* alive, int& idle, unsigned ns)
{
  __nanosleep(ns);
  if (++idle > MCW_PUMP_IDLE_LIMIT && (threadIdx.x & 31) == 0)
    atomicExch(reinterpret_cast<unsigned long long*>(alive), (unsigned long long)(-(1ll << 60)));
}

// Warp-convergent push: lanes with `want` append history i.  The caller has written the history's state.
MCW_D void mcw_ring_push(int32_t* ring, unsigned long long* tail, uint32_t mask, bool want, int i)
{
  if (!__any_sync(MCW_FULL, want))
    return;
  __threadfence(); // release: the state of the history before its index
  const unsigned long long t = mcw_warp_append(tail, want);
  if (want)
    *reinterpret_cast<volatile int32_t*>(ring + (t & mask)) = i;
}

// Non-blocking take of the slot of `ticket`; -1 if the producer has not arrived yet.
MCW_D int mcw_ring_poll(int32_t* ring, uint32_t mask, unsigned long long ticket)
{
  int32_t* s = ring + (ticket & mask);
  const int v = *reinterpret_cast<volatile int32_t*>(s);
  if (v >= 0)
    *reinterpret_cast<volatile int32_t*>(s) = -1; // the slot is reused a whole lap (>= 2 n tickets) later
  return v;
}

// ------------------------------------------------------------------ tracking warps
template <int TYPE>
MCW_D void mcw_pump_track(const McwParams& P, const McwPump& Q)
{
  const unsigned lane = threadIdx.x & 31u;
  const McwParticlesDev& p = P.p;
  const bool track_tally = P.tallies_active && P.any_tracklength;
  const bool keep_hash = P.record_traces != 0;
  enum
  {
    LN_EMPTY = 0,
    LN_WAIT = 4, // holds a ticket of the fly ring
    LN_PUSH = 5, // state written back, waits to be appended to the ring `dest` (appends are batched per warp)
    LN_TRACK = MCW_LN_TRACK,
    LN_REFLECT = MCW_LN_REFLECT
  };
  int state = LN_EMPTY, dest = 0; // dest: 0 collide ring, 1 general-walk ring, 2 fly ring (visit cap), 3 dead
  unsigned long long ticket = 0;
  int i = -1;
  V3 point = v3(0, 0, 0), dir = v3(0, 0, 1);
  int cur = 0, last = -1, inc_side = -1, slot = -1, n_seg = 0, visit = 0;
  SegState s;
  s.E = s.wgt = s.sig_a = s.sig_f = s.sig_nf = s.k_trk = 0.0;
  s.sig_t = 1.0;
  s.seed_trk = s.seed_src = s.seed_urr = 0;
  uint64_t hash = 0;
  unsigned segs = 0;
  double hmax_c = 0.0, inv_hmax = 0.0;
  int idle = 0;

  for (;;)
  {
    if (__shfl_sync(MCW_FULL, mcw_ld_alive(Q.alive), 0) <= 0)
      break;
    // idle lanes take a ticket of the fly ring (a few at a time: one atomic per warp) and poll their slot
    {
      const bool want = state == LN_EMPTY;
      const unsigned m_want = __ballot_sync(MCW_FULL, want);
      if (__popc(m_want) >= MCW_REFILL_MIN || (m_want != 0u && !__any_sync(MCW_FULL, state == LN_TRACK)))
      {
        const unsigned long long t = mcw_warp_append(Q.fly_head, want);
        if (want)
        {
          ticket = t;
          state = LN_WAIT;
        }
      }
    }
    int got = -1;
    if (state == LN_WAIT)
      got = mcw_ring_poll(Q.q_fly, Q.mask, ticket);
    if (__any_sync(MCW_FULL, got >= 0))
    {
      __threadfence(); // acquire: the producer's state writes
      if (got >= 0)
      {
        i = got;
        if (__ldcg(p.ext_first + i) != MCW_INVALID)
        {
          // last hit was a vertex / edge: the general walk takes the next step (RT TraceRay.C:1134)
          p.status[i] = MCW_ST_FALLBACK;
          dest = 1;
          state = LN_PUSH;
        }
        else
        {
          point = v3(__ldcg(p.x + i), __ldcg(p.y + i), __ldcg(p.z + i));
          dir = v3(__ldcg(p.ux + i), __ldcg(p.uy + i), __ldcg(p.uz + i));
          cur = __ldcg(p.elem + i);
          last = __ldcg(p.last_elem + i);
          inc_side = __ldcg(p.inc_side + i);
          s.E = __ldcg(p.E + i);
          s.wgt = __ldcg(p.wgt + i);
          s.seed_trk = __ldcg(p.seed_trk + i);
          s.k_trk = __ldcg(p.k_trk + i);
          n_seg = __ldcg(p.n_seg + i);
          if (keep_hash)
            hash = __ldcg(p.hash + i);
          slot = -1;
          visit = 0;
          state = LN_TRACK;
        }
      }
    }

    const unsigned m_refl = __ballot_sync(MCW_FULL, state == LN_REFLECT);
    const unsigned m_track = __ballot_sync(MCW_FULL, state == LN_TRACK);
    if (m_refl != 0u && (__popc(m_refl) >= MCW_REFLECT_MIN || m_track == 0u))
    {
      if (state == LN_REFLECT)
      {
        // RT ReflectRayBC.C:68-76 + Ray::changeDirection (Ray.C:206-214); inc_side holds the side hit
        ElemRec<TYPE> R(P.mesh, cur);
        V3 n = mcw_side_normal<TYPE>(R, inc_side);
        double dn = dot(dir, n);
        V3 r = dir - (2.0 * dn) * n;
        r = r / norm(r);
        dir = unit(r);
        last = cur;
        state = LN_TRACK;
      }
    }
    // histories that left their lane go to their next stage, a few per append; at once when nothing else moves
    {
      const unsigned m_push = __ballot_sync(MCW_FULL, state == LN_PUSH);
      if (m_push != 0u && (__popc(m_push) >= MCW_REFILL_MIN || m_track == 0u))
      {
        const bool mine = state == LN_PUSH;
        mcw_ring_push(Q.q_collide, Q.collide_tail, Q.mask, mine && dest == 0, i);
        mcw_ring_push(Q.q_fallback, Q.fallback_tail, Q.mask, mine && dest == 1, i);
        mcw_ring_push(Q.q_fly, Q.fly_tail, Q.mask, mine && dest == 2, i);
        const unsigned m_dead = __ballot_sync(MCW_FULL, mine && dest == 3);
        if (lane == 0 && m_dead)
          atomicAdd(reinterpret_cast<unsigned long long*>(Q.alive), (unsigned long long)(-(long long)__popc(m_dead)));
        if (mine)
          state = LN_EMPTY;
      }
    }
    if ((m_refl | m_track) == 0u)
    {
      mcw_pump_idle(Q.alive, idle, 200); // nothing to walk: wait for the rings
      continue;
    }
    idle = 0;

    if (state == LN_TRACK)
    {
      int status = MCW_ST_FLY, fate = -1;
      bool leave = false, requeue = false;
      ElemData<TYPE> D;
      D.load(P.mesh, cur);
      uint64_t seed_next = s.seed_trk;
      const double neg_log = -mcw_log(mcw_prn(seed_next));
      if (D.hmax != hmax_c)
      {
        hmax_c = D.hmax;
        inv_hmax = 1.0 / hmax_c;
      }
      if (D.slot != slot)
      {
        slot = D.slot;
        const double2* c = reinterpret_cast<const double2*>(p.xs + ((size_t)i * P.n_slots + slot) * 4);
        const double2 a = __ldcg(c), b = __ldcg(c + 1);
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
      else if (visit >= P.visit_cap)
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
          status = MCW_ST_FALLBACK; // nothing modified: the general walk redoes this step
          leave = true;
        }
        else
        {
          ++n_seg;
          ++segs;
          ++visit;
          V3 pos;
          const int code = mcw_pick(D.nbr, iside);
          const int elem = cur;
          s.seed_trk = seed_next;
          const bool collided = mcw_segment_nl(P, s, neg_log, elem, D.block, P.slot_mat[slot], ipoint, dir, idist, pos);
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
              last = cur;
              cur = code >> MCW_NBR_SHIFT;
              inc_side = code & ((1 << MCW_NBR_SHIFT) - 1);
            }
            else
            {
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
                state = LN_REFLECT;
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
        dest = requeue ? 2 : (status == MCW_ST_COLLIDE ? 0 : (status == MCW_ST_FALLBACK ? 1 : 3));
        state = LN_PUSH;
      }
    }
  }
  {
    unsigned tot = segs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      tot += __shfl_xor_sync(MCW_FULL, tot, o);
    if (lane == 0 && tot)
      atomicAdd(&P.counters->n_segments, (unsigned long long)tot);
  }
}

// ------------------------------------------------------------------ collision warps
// A warp holds 32 tickets of the collide ring; when all have arrived (or `patience` polls have passed since
// the first did) it runs event_collide for the arrived ones, refreshes the survivors' cross sections and
// returns them to the fly ring.
template <int TYPE>
MCW_D void mcw_pump_collide(const McwParams& P, const McwPump& Q, int patience)
{
  const unsigned lane = threadIdx.x & 31u;
  bool has_ticket = false, have = false;
  unsigned long long ticket = 0;
  int i = 0, waited = 0, idle = 0;
  for (;;)
  {
    if (__shfl_sync(MCW_FULL, mcw_ld_alive(Q.alive), 0) <= 0)
      break;
    {
      const bool want = !has_ticket && !have;
      if (__any_sync(MCW_FULL, want))
      {
        const unsigned long long t = mcw_warp_append(Q.collide_head, want);
        if (want)
        {
          ticket = t;
          has_ticket = true;
        }
      }
    }
    if (has_ticket)
    {
      const int v = mcw_ring_poll(Q.q_collide, Q.mask, ticket);
      if (v >= 0)
      {
        i = v;
        have = true;
        has_ticket = false;
      }
    }
    const unsigned m_have = __ballot_sync(MCW_FULL, have);
    if (m_have == 0u)
    {
      mcw_pump_idle(Q.alive, idle, 200);
      continue;
    }
    idle = 0;
    if (m_have != MCW_FULL && ++waited < patience)
      continue;
    waited = 0;
    __threadfence(); // acquire: the tracking warps' state writes
    const bool survived = mcw_collide_history<TYPE, false>(P, have ? i : 0, have) && have;
    const unsigned m_surv = __ballot_sync(MCW_FULL, survived);
    // Particle::event_calculate_xs at the survivors' new energies
    if (__popc(m_surv) >= MCW_PUMP_COOP_MAX)
    {
      if (survived)
        mcw_lookup_all_slots(P, i, P.p.E[i]);
    }
    else
      mcw_warp_lookup_all_slots(P, m_surv, i, survived ? P.p.E[i] : 0.0);
    mcw_ring_push(Q.q_fly, Q.fly_tail, Q.mask, survived, i);
    const int n_dead = __popc(m_have & ~m_surv);
    if (lane == 0 && n_dead)
      atomicAdd(reinterpret_cast<unsigned long long*>(Q.alive), (unsigned long long)(-(long long)n_dead));
    have = false;
  }
}

// ------------------------------------------------------------------ general-walk warps (rare path)
template <int TYPE>
MCW_D void mcw_pump_fallback(const McwParams& P, const McwPump& Q)
{
  const unsigned lane = threadIdx.x & 31u;
  bool has_ticket = false;
  unsigned long long ticket = 0;
  int idle = 0;
  for (;;)
  {
    if (__shfl_sync(MCW_FULL, mcw_ld_alive(Q.alive), 0) <= 0)
      break;
    {
      const bool want = !has_ticket;
      if (__any_sync(MCW_FULL, want))
      {
        const unsigned long long t = mcw_warp_append(Q.fallback_head, want);
        if (want)
        {
          ticket = t;
          has_ticket = true;
        }
      }
    }
    const int v = mcw_ring_poll(Q.q_fallback, Q.mask, ticket);
    if (!__any_sync(MCW_FULL, v >= 0))
    {
      mcw_pump_idle(Q.alive, idle, 500);
      continue;
    }
    idle = 0;
    __threadfence();
    int status = -1;
    if (v >= 0)
    {
      has_ticket = false;
      status = mcw_fallback_history<TYPE>(P, v);
    }
    mcw_ring_push(Q.q_fly, Q.fly_tail, Q.mask, status == MCW_ST_FLY, v);
    mcw_ring_push(Q.q_collide, Q.collide_tail, Q.mask, status == MCW_ST_COLLIDE, v);
    const unsigned m_dead = __ballot_sync(MCW_FULL, status == MCW_ST_DEAD);
    if (lane == 0 && m_dead)
      atomicAdd(reinterpret_cast<unsigned long long*>(Q.alive), (unsigned long long)(-(long long)__popc(m_dead)));
  }
}

template <int TYPE>
__global__ void __launch_bounds__(MCW_TRACK_BLOCK, MCW_PUMP_MIN_BLOCKS)
k_pump(McwParams P, McwPump Q, int n_track_warps, int patience)
{
  const int warp = threadIdx.x >> 5;
  if (blockIdx.x < MCW_PUMP_FB_BLOCKS && warp == (int)(blockDim.x >> 5) - 1)
    mcw_pump_fallback<TYPE>(P, Q);
  else if (warp < n_track_warps)
    mcw_pump_track<TYPE>(P, Q);
  else
    mcw_pump_collide<TYPE>(P, Q, patience);
}

// Fills the rings from the event loop's queues: the fly queue (status MCW_ST_FLY) and the histories
// waiting for a general step (status MCW_ST_FALLBACK).  Entries that no longer have that status are skipped.
__global__ void k_pump_fill(McwPump Q, const int8_t* __restrict__ status, const int32_t* __restrict__ q_fly,
                            const unsigned long long* __restrict__ n_fly, const int32_t* __restrict__ q_fb,
                            const unsigned long long* __restrict__ n_fb)
{
  const int64_t nf = (int64_t)*n_fly, nb = (int64_t)*n_fb;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nf + nb; idx += (int64_t)gridDim.x * blockDim.x)
  {
    const bool fly = idx < nf;
    const int i = fly ? q_fly[idx] : q_fb[idx - nf];
    if (status[i] != (fly ? MCW_ST_FLY : MCW_ST_FALLBACK))
      continue;
    const unsigned long long t = atomicAdd(fly ? Q.fly_tail : Q.fallback_tail, 1ull);
    (fly ? Q.q_fly : Q.q_fallback)[t & Q.mask] = i;
    atomicAdd(reinterpret_cast<unsigned long long*>(Q.alive), 1ull);
  }
}
