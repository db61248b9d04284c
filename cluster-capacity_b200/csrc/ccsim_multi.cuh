// ccsim_multi.cuh — multi-commit waves for templates coupled through per-domain counters (PodTopologySpread DoNotSchedule,
// required pod anti-affinity): several reference scheduling cycles per grid-wide exchange, identical pod -> node sequence.
//
// Why it is legal. Within a stretch of cycles in which no PodTopologySpread global minimum changes, a node's feasibility
// depends on its own row and on the match counts of its topology domains only, and it is MONOTONE: counts only grow
// (inc >= 0), so a feasible node can become infeasible but never the reverse; a node's score changes only when the node
// itself is committed. The winner of each cycle is therefore the highest-keyed node that is still feasible under the
// counts as updated so far (podtopologyspread/filtering.go:311-356, interpodaffinity/filtering.go:352-432;
// schedule_one.go:894-941 with the canonical first-max tie rule). If every CTA publishes its M best feasible nodes (the
// exact top-M: keys are unique) together with the domain ids their feasibility depends on, every CTA can replay those
// cycles redundantly and deterministically: walk all published candidates in descending key order, re-check the per-domain
// terms against its own replicated counters, commit the ones that still pass (candidates that fail are dropped from all
// lists in parallel: by monotonicity they cannot come back during the wave). The replay stops when
//   (a) a PTS minimum moves (limits change, rejected nodes may come back: rescan),
//   (b) a CTA with more feasible nodes than it published has used up its list and the next candidate's key is below that
//       CTA's last published key (an unpublished node could rank in between),
//   (c) the pod limit is reached (simulator.go:300-305).
// The first candidate of a wave is always accepted, so every wave makes progress; a wave without candidates is the
// Unschedulable stop. Node-local terms (hostname anti-affinity, ...) need no re-check: a node appears once per wave.
//
// Exchange: one 128-byte line per CTA = M (key, payload) pairs, each word tagged (self-validating, no fences). All 24 warps
// gather the 148 lines in one L2 round trip; warp 0 of every CTA replays.
#pragma once
#include "ccsim_lean.cuh"

#define MULTI_M 8                 /* candidates per CTA and wave: 8 x 16 B = the CTA's 128-byte slot line */
#define MULTI_PAY_BITS 39         /* payload bits for domain ids (dom+1 per topology slot) */
#define MULTI_MORE_BIT 39
#define MULTI_LEN_SHIFT 40

struct __align__(16) MultiShared {
  unsigned long long wtop[LEAN_WARPS][MULTI_M];     // per-warp top-M keys of this wave
  unsigned long long gkey[CCSIM_MAX_GRID][MULTI_M]; // gathered candidate keys (tag stripped) ...
  unsigned long long gpay[CCSIM_MAX_GRID][MULTI_M]; // ... and payloads
  int32_t wfeas[LEAN_WARPS];
  int32_t accepted, dead, pad[2];
};

__shared__ MultiShared ms;

struct MultiParams {
  uint32_t pay_shift[LEAN_MAX_SLOTS];   // record slot s (a topology column) -> bit position of its dom+1 field in the payload
  uint32_t pay_mask[LEAN_MAX_SLOTS];    // field mask (0: the slot is a node-local counter, not carried)
};

__device__ __forceinline__ void ld_slot2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__global__ void __launch_bounds__(LEAN_THREADS, 1) ccsim_wave_multi_kernel(const DevParams p, const LeanParams lp, const MultiParams mp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *smem_cnt = reinterpret_cast<int32_t *>(smem_raw);
  const uint32_t cnt_bytes = ((uint32_t)p.smem_cnt_ints * 4u + 15u) & ~15u;
  uint4 *rec = reinterpret_cast<uint4 *>(smem_raw + cnt_bytes);
  const size_t cp = (size_t)p.chunk_pad;
  long long *c_acpu = reinterpret_cast<long long *>(smem_raw + cnt_bytes + lp.rec_bytes_total);
  long long *c_amem = c_acpu + cp, *c_rcpu = c_amem + cp, *c_rmem = c_rcpu + cp, *c_zcpu = c_rmem + cp, *c_zmem = c_zcpu + cp;
  int32_t *c_apods = reinterpret_cast<int32_t *>(c_zmem + cp);
  int32_t *c_npods = c_apods + cp;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int32_t lo = min(p.n, cta * p.chunk), hi = min(p.n, lo + p.chunk);
  const int32_t cnt_nodes = hi - lo;      // <= LEAN_THREADS (host-checked): one node per thread
  const int su = lp.stride_u;

  // ---- stage the tile (once): hot AoS records + cold SoA columns (same layout as the lean kernel) ----
  for (int32_t j = tid; j < cnt_nodes; j += LEAN_THREADS) {
    const int32_t i = lo + j;
    const long long ac = p.alloc_cpu[i], am = p.alloc_mem[i], rc = p.req_cpu[i], rm = p.req_mem[i];
    const int32_t ap = p.alloc_pods[i], np = p.npods[i];
    unsigned long long *r8 = reinterpret_cast<unsigned long long *>(rec + (size_t)j * su);
    int32_t *r4 = reinterpret_cast<int32_t *>(r8);
    r8[0] = p.taint_mask[i];
    r8[1] = p.static_words > 0 ? p.static_mask[i] : 0ull;
    r8[2] = (unsigned long long)(ac - rc);
    r8[3] = (unsigned long long)(am - rm);
    r4[8] = ap - np;
    r4[9] = -1;
    for (int s = 0; s < lp.n_slots; s++)
      r4[10 + s] = lp.slot_topo[s] >= 0 ? p.topo[lp.slot_topo[s]][i] : p.counters[lp.slot_counter[s]].work[i];
    c_acpu[j] = ac; c_amem[j] = am; c_rcpu[j] = rc; c_rmem[j] = rm;
    c_zcpu[j] = p.nz_cpu[i]; c_zmem[j] = p.nz_mem[i];
    c_apods[j] = ap; c_npods[j] = np;
  }
  for (int k = tid; k < (int)(sizeof(ccsim_template) / 8); k += LEAN_THREADS)
    reinterpret_cast<unsigned long long *>(&ls.tmpl)[k] = reinterpret_cast<const unsigned long long *>(&p.templates[0])[k];
  for (int j = 0; j < p.n_counters; j++) {
    const DevCounter &dc = p.counters[j];
    if (dc.topo_col < 0) continue;
    for (int d = tid; d < dc.n_domains; d += LEAN_THREADS) smem_cnt[dc.smem_off + d] = dc.init[d];
  }
  if (tid == 0) { ls.aff_total = p.templates[0].aff_total_init; ls.winner = -1; ls.stop = 0; ls.dirty = 1; ms.accepted = 0; ms.dead = 0; }
  __syncthreads();
  for (int c = 0; c < ls.tmpl.n_pts; c++) lean_pts_recount(p, smem_cnt, c);

  long long k = 0, wv = 0;
  uint32_t wtag = 1;
  uint32_t tag = (p.epoch << 12) | wtag;
  for (;; wv++) {
    if (p.max_pods > 0 && k >= p.max_pods) { if (tid == 0) ls.stop = 2; __syncthreads(); break; }
    if (k > p.pod_cap) { if (tid == 0) ls.stop = 3; __syncthreads(); break; }
    if (ls.dirty) {
      if (tid == 0) lean_build_consts(p, lp);
      __syncthreads();
      if (tid == 0) ls.dirty = 0;
    }
    // ---- fused Filter pass: this thread's node ----
    unsigned long long key = 0ull;
    if (tid < cnt_nodes) {
      const int32_t j = tid;
      const uint4 *r = rec + (size_t)j * su;
      const uint4 u0 = r[0], u1 = r[1], u2 = r[2];
      const unsigned long long taint0 = ((unsigned long long)u0.y << 32) | u0.x;
      const unsigned long long static0 = ((unsigned long long)u0.w << 32) | u0.z;
      const long long free_cpu = (long long)(((unsigned long long)u1.y << 32) | u1.x);
      const long long free_mem = (long long)(((unsigned long long)u1.w << 32) | u1.z);
      const int32_t free_pods = (int32_t)u2.x;
      int32_t sc = (int32_t)u2.y;
      bool ok = ((taint0 & ls.taint_bad0) | (~static0 & ls.sel0) | (static0 & ls.forbid0)) == 0ull;
      ok &= (free_cpu >= ls.eq_cpu) & (free_mem >= ls.eq_mem) & (free_pods >= ls.pods_need);
      const int32_t n_cmp = ls.n_cmp_terms;
      const int32_t *r4 = reinterpret_cast<const int32_t *>(r);
      #pragma unroll 4
      for (int q = 0; q < n_cmp; q++) {
        const LeanTerm lt = ls.terms[q];
        const int32_t v = r4[lt.slot];
        const bool local = lt.cnt_off < 0;
        const int32_t c = local ? v : smem_cnt[lt.cnt_off + (v < 0 ? 0 : v)];
        const bool has = local | (v >= 0);
        ok &= has ? (c <= lt.lim) : (lt.miss_rejects == 0);
      }
      if (ok) {
        if (sc < 0) {
          sc = score_node(c_acpu[j], c_amem[j], c_zcpu[j] + ls.tmpl.least_cpu, c_zmem[j] + ls.tmpl.least_mem,
                          c_rcpu[j] + ls.tmpl.bal_cpu, c_rmem[j] + ls.tmpl.bal_mem, ls.sw);
          reinterpret_cast<int32_t *>(rec + (size_t)j * su)[9] = sc;
        }
        key = pack_key(sc, (uint32_t)(p.node_base + lo + j));
      }
    }
    // ---- the warp's M best keys (REDUX rounds; keys are unique, 0 = none) ----
    {
      unsigned long long rem = key;
      const int nf = __popc(__ballot_sync(0xffffffffu, key != 0ull));
      if (lane == 0) ms.wfeas[warp] = nf;
      #pragma unroll
      for (int r = 0; r < MULTI_M; r++) {
        unsigned long long v = 0ull;
        if (r < nf) {                         // warp-uniform: no REDUX rounds for entries that do not exist
          v = warp_max_u64(rem);
          if (rem == v) rem = 0ull;
        }
        if (lane == 0) ms.wtop[warp][r] = v;
      }
    }
    __syncthreads();                                                    // S1
    const unsigned long long tagbits = (unsigned long long)tag << KEY_TAG_SHIFT;
    if (warp == 0) {
      // ---- the CTA's M best: merge of the 24 sorted warp lists, then publish (key, payload) pairs ----
      unsigned long long e[MULTI_M];
      #pragma unroll
      for (int r = 0; r < MULTI_M; r++) e[r] = lane < LEAN_WARPS ? ms.wtop[lane][r] : 0ull;
      int32_t total = lane < LEAN_WARPS ? ms.wfeas[lane] : 0;
      total = __reduce_add_sync(0xffffffffu, total);
      int ptr = 0, L = 0;
      unsigned long long mykey = 0ull;
      #pragma unroll
      for (int r = 0; r < MULTI_M; r++) {
        unsigned long long head = 0ull;
        #pragma unroll
        for (int q = 0; q < MULTI_M; q++) if (q == ptr) head = e[q];
        const unsigned long long g = warp_max_u64(head);
        if (g != 0ull) {
          if (head == g) ptr++;
          if (lane == r) mykey = g;
          L = r + 1;
        }
      }
      unsigned long long pay = 0ull;
      if (lane < L) {
        const int32_t jj = (int32_t)key_index(mykey) - (p.node_base + lo);
        const int32_t *r4 = reinterpret_cast<const int32_t *>(rec + (size_t)jj * su);
        for (int s = 0; s < lp.n_slots; s++)
          if (mp.pay_mask[s]) pay |= (unsigned long long)((uint32_t)(r4[10 + s] + 1) & mp.pay_mask[s]) << mp.pay_shift[s];
      }
      pay |= ((unsigned long long)L << MULTI_LEN_SHIFT) | ((unsigned long long)(total > L ? 1 : 0) << MULTI_MORE_BIT);
      unsigned long long *myslots = p.slots + ((size_t)(wv & 1) * CCSIM_MAX_GRID + cta) * SLOT_STRIDE;
      if (lane < MULTI_M) {
        st_slot(&myslots[2 * lane], (mykey & KEY_BODY_MASK) | tagbits);
        st_slot(&myslots[2 * lane + 1], pay | tagbits);
      }
    }
    // ---- gather: every (CTA, entry) pair is polled by one thread; all loads of the CTA are in flight together ----
    {
      const unsigned long long *base = p.slots + (size_t)(wv & 1) * CCSIM_MAX_GRID * SLOT_STRIDE;
      const int tot = p.grid * MULTI_M;
      const int e0 = tid, e1 = tid + LEAN_THREADS;
      bool need0 = e0 < tot, need1 = e1 < tot;
      unsigned long long a0 = 0, b0 = 0, a1 = 0, b1 = 0;
      unsigned spins = 0;
      while (need0 | need1) {
        if (need0) ld_slot2(base + (size_t)(e0 >> 3) * SLOT_STRIDE + 2 * (e0 & 7), a0, b0);
        if (need1) ld_slot2(base + (size_t)(e1 >> 3) * SLOT_STRIDE + 2 * (e1 & 7), a1, b1);
        if (need0 && (uint32_t)(a0 >> KEY_TAG_SHIFT) == tag && (uint32_t)(b0 >> KEY_TAG_SHIFT) == tag) {
          need0 = false; ms.gkey[e0 >> 3][e0 & 7] = a0 & KEY_BODY_MASK; ms.gpay[e0 >> 3][e0 & 7] = b0 & KEY_BODY_MASK;
        }
        if (need1 && (uint32_t)(a1 >> KEY_TAG_SHIFT) == tag && (uint32_t)(b1 >> KEY_TAG_SHIFT) == tag) {
          need1 = false; ms.gkey[e1 >> 3][e1 & 7] = a1 & KEY_BODY_MASK; ms.gpay[e1 >> 3][e1 & 7] = b1 & KEY_BODY_MASK;
        }
        if (++spins > WATCHDOG_SPINS) { ms.dead = 1; break; }
      }
    }
    __syncthreads();                                                    // S2
    if (warp == 0) {
      // ---- replay: the reference cycles k, k+1, ... this wave can decide (identical in every CTA) ----
      const ccsim_template &t = ls.tmpl;
      unsigned long long headk[CCSIM_MAX_GRID / 32];
      int32_t ptrs[CCSIM_MAX_GRID / 32], lens[CCSIM_MAX_GRID / 32], mores[CCSIM_MAX_GRID / 32];
      #pragma unroll
      for (int q = 0; q < CCSIM_MAX_GRID / 32; q++) {
        const int c = lane + 32 * q;
        ptrs[q] = 0; lens[q] = 0; mores[q] = 0; headk[q] = 0ull;
        if (c < p.grid) {
          const unsigned long long meta = ms.gpay[c][0];
          lens[q] = (int32_t)((meta >> MULTI_LEN_SHIFT) & 15ull);
          mores[q] = (int32_t)((meta >> MULTI_MORE_BIT) & 1ull);
          headk[q] = lens[q] > 0 ? ms.gkey[c][0] : 0ull;
        }
      }
      const int32_t n_cmp = ls.n_cmp_terms;
      int32_t acc = 0;
      const bool dead = ms.dead != 0;
      unsigned long long barrier = 0ull;      // highest "last published key" of a CTA whose list ran out while it has more nodes
      while (!dead) {
        // prune: a candidate that is infeasible under the current counts stays infeasible for the rest of the wave
        // (monotone), so every lane drops such heads of its own lists at once, whatever their rank
        #pragma unroll
        for (int q = 0; q < CCSIM_MAX_GRID / 32; q++) {
          const int c = lane + 32 * q;
          while (ptrs[q] < lens[q]) {
            const unsigned long long py = ms.gpay[c][ptrs[q]];
            bool bad = false;
            for (int tq = 0; tq < n_cmp; tq++) {
              const LeanTerm lt = ls.terms[tq];
              if (lt.cnt_off < 0) continue;
              const int sl = lt.slot - 10;
              const int32_t v = (int32_t)((uint32_t)(py >> mp.pay_shift[sl]) & mp.pay_mask[sl]) - 1;
              const bool has = v >= 0;
              const int32_t cv = smem_cnt[lt.cnt_off + (has ? v : 0)];
              bad |= has ? (cv > lt.lim) : (lt.miss_rejects != 0);
            }
            if (!bad) break;
            ptrs[q]++;
          }
          if (ptrs[q] < lens[q]) headk[q] = ms.gkey[c][ptrs[q]];
          else {
            headk[q] = 0ull;
            if (mores[q] && lens[q] > 0) { const unsigned long long lastk = ms.gkey[c][lens[q] - 1]; barrier = lastk > barrier ? lastk : barrier; }
          }
        }
        unsigned long long h = 0ull; int hq = 0;
        #pragma unroll
        for (int q = 0; q < CCSIM_MAX_GRID / 32; q++) if (headk[q] > h) { h = headk[q]; hq = q; }
        const unsigned long long g = warp_max_u64(h);
        if (g == 0ull) break;                                   // every published list is used up
        // nodes a CTA did not publish have smaller keys than its last published one: g is the true maximum only above that
        if (g < warp_max_u64(barrier)) break;
        const int ol = __ffs(__ballot_sync(0xffffffffu, h == g)) - 1;
        unsigned long long pay = 0ull;
        if (lane == ol) {
          const int c = lane + 32 * hq;
          #pragma unroll
          for (int q = 0; q < CCSIM_MAX_GRID / 32; q++) if (q == hq) { pay = ms.gpay[c][ptrs[q]]; ptrs[q]++; }
        }
        pay = __shfl_sync(0xffffffffu, pay, ol);
        bool stop_batch = false;
        {
          // ---- commit pod k+acc (assume -> AssumePod -> NodeInfo.update(+1): schedule_one.go:967-984, types.go:409-427) ----
          const int32_t gi = (int32_t)key_index(g);
          const int32_t w = gi - p.node_base;
          const bool mine = (w >= lo && w < hi);
          const int32_t jw = w - lo;
          if (mine && lane == 31) {
            const long long rc = c_rcpu[jw] + t.req_cpu, rm = c_rmem[jw] + t.req_mem;
            const long long zc = c_zcpu[jw] + t.nz_cpu, zm = c_zmem[jw] + t.nz_mem;
            const int32_t np = c_npods[jw] + 1;
            c_rcpu[jw] = rc; c_rmem[jw] = rm; c_zcpu[jw] = zc; c_zmem[jw] = zm; c_npods[jw] = np;
            unsigned long long *r8 = reinterpret_cast<unsigned long long *>(rec + (size_t)jw * su);
            int32_t *r4 = reinterpret_cast<int32_t *>(r8);
            r8[2] = (unsigned long long)(c_acpu[jw] - rc);
            r8[3] = (unsigned long long)(c_amem[jw] - rm);
            r4[8] = c_apods[jw] - np;
            r4[9] = -1;
            p.req_cpu[w] = rc; p.req_mem[w] = rm; p.nz_cpu[w] = zc; p.nz_mem[w] = zm; p.npods[w] = np;   // write through
            if (k + acc < p.pod_cap) p.pod_node[k + acc] = gi; else ls.stop = 3;
          }
          bool minchg = false;
          if (lane < p.n_counters) {
            const int j = lane;
            const CommitInfo ci = ls.cinfo[j];
            if (ci.inc) {
              const int s = lp.counter_slot[j];
              if (ci.local) {
                if (mine) {
                  int32_t *r4 = reinterpret_cast<int32_t *>(rec + (size_t)jw * su);
                  const int32_t nv = r4[10 + s] + ci.inc;
                  r4[10 + s] = nv;
                  p.counters[j].work[w] = nv;
                }
              } else {
                const int32_t dom = (int32_t)((uint32_t)(pay >> mp.pay_shift[s]) & mp.pay_mask[s]) - 1;
                if (dom >= 0) {
                  int32_t *cnt = smem_cnt + p.counters[j].smem_off;
                  const int32_t old = cnt[dom];
                  cnt[dom] = old + ci.inc;
                  if (ci.pts_idx >= 0 && dom < ci.n_present && old == ls.ptsmin[ci.pts_idx]) {
                    const int32_t left = ls.ptsnum[ci.pts_idx] - 1;
                    ls.ptsnum[ci.pts_idx] = left;
                    minchg = left <= 0;          // the global minimum of this constraint moves: limits change, rescan
                  }
                }
              }
            }
          }
          acc++;
          stop_batch |= __any_sync(0xffffffffu, minchg);
          stop_batch |= (p.max_pods > 0 && k + acc >= p.max_pods);
          stop_batch |= (k + acc >= p.pod_cap);
        }
        __syncwarp();
        if (stop_batch) break;
      }
      // a PTS minimum moved: recount it here (warp 0 owns the counters during the replay) and move the term's limit, instead
      // of a block-wide recount + rebuild of all Filter constants (filtering.go:56-69,98-137: minMatchNum / criticalPaths)
      __syncwarp();
      for (int c = 0; c < t.n_pts; c++) {
        if (t.pts[c].min_zero || ls.ptsnum[c] > 0) continue;       // uniform: shared memory, written before the __syncwarp
        const DevCounter &dc = p.counters[t.pts[c].counter];
        if (dc.n_present <= 0) continue;
        const int32_t *cnt = smem_cnt + dc.smem_off;
        int32_t m = INT32_MAX;
        for (int d = lane; d < dc.n_present; d += 32) m = min(m, cnt[d]);
        m = __reduce_min_sync(0xffffffffu, m);
        int32_t num = 0;
        for (int d = lane; d < dc.n_present; d += 32) num += (cnt[d] == m);
        num = __reduce_add_sync(0xffffffffu, num);
        if (lane == 0) {
          ls.ptsmin[c] = m; ls.ptsnum[c] = num;
          if (t.filter_enable & CCSIM_PL_POD_TOPOLOGY_SPREAD) {    // terms[0..n_pts) are the PTS terms, in constraint order
            const long long lim = (long long)t.pts[c].max_skew - t.pts[c].self_match + (long long)m;
            ls.terms[c].lim = lim > INT32_MAX ? INT32_MAX : (lim < INT32_MIN ? INT32_MIN : (int32_t)lim);
          }
        }
        __syncwarp();
      }
      if (lane == 0) {
        ms.accepted = acc;
        if (dead) ls.stop = 3;
        else if (acc == 0) ls.stop = 1;          // no feasible node anywhere: the pod is unschedulable
      }
    }
    __syncthreads();                                                    // S3
    k += ms.accepted;
    if (ls.stop) break;
    for (int c = 0; c < ls.tmpl.n_pts; c++)
      if (!ls.tmpl.pts[c].min_zero && ls.ptsnum[c] <= 0 && p.counters[ls.tmpl.pts[c].counter].n_present > 0) lean_pts_recount(p, smem_cnt, c);
    wtag = (wtag == 4095u) ? 1u : wtag + 1u;
    tag = (p.epoch << 12) | wtag;
  }

  if (cta == 0) {
    for (int j = 0; j < p.n_counters; j++) {
      const DevCounter &dc = p.counters[j];
      if (dc.topo_col < 0) continue;
      for (int d = tid; d < dc.n_domains; d += LEAN_THREADS) p.final_cnt[p.final_off[j] + d] = smem_cnt[dc.smem_off + d];
    }
    if (tid == 0) {
      DevOut *o = p.out;
      o->placed = k;
      o->stop_code = (ls.stop == 2) ? CCSIM_STOP_LIMIT_REACHED : CCSIM_STOP_UNSCHEDULABLE;
      o->error = (ls.stop == 3) ? 1 : 0;
      o->waves = (ls.stop == 2) ? wv : wv + 1;
      o->evals = o->waves * (long long)p.n;
      o->examined = o->evals;
      for (int c = 0; c < CCSIM_MAX_PTS; c++) o->ptsmin[c] = ls.ptsmin[c];
      o->aff_total = ls.aff_total;
    }
  }
}
