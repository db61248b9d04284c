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
// A committed node may win again inside the wave with its new score, which is in nobody's list (a randomized differential test,
// tests/test_gpu_stress.py, found this hole in the first version: spread-only templates). Every published candidate therefore
// carries, next to its domain ids, its key after one more clone (node-local Filter part + score recomputed by the publisher; 0 =
// it would not fit). The winner comes back into the replay once with that key ("second life"); when a node wins for the second
// time in a wave its third key is unknown and the wave ends after that commit. Entries in their second life do not keep a list
// "alive" for rule (b): the CTA's unpublished nodes rank below its last PUBLISHED key only.
// The first candidate of a wave is always accepted, so every wave makes progress; a wave without candidates is the
// Unschedulable stop. Node-local terms (hostname anti-affinity, ...) need no re-check: a node appears once per wave.
//
// Exchange: one 128-byte line per CTA = M (key, payload) pairs, each word tagged (self-validating, no fences). All 24 warps
// gather the 148 lines in one L2 round trip; warp 0 of every CTA replays.
#pragma once
#include "ccsim_lean.cuh"

#define MULTI_M 8                 /* candidates per CTA and wave: 8 x 16 B = the CTA's 128-byte slot line */
#define MULTI_PAY_BITS 27         /* payload bits for domain ids (dom+1 per topology slot) */
#define MULTI_NEXT_SHIFT 27       /* 12 bits: (score + 1) of the node after one more clone, 0 = it would not fit any more */
#define MULTI_MORE_BIT 39
#define MULTI_LEN_SHIFT 40
#define MULTI_MAX_ACC 64          /* commits one wave may decide */
#define MULTI_GT 6                /* Filter terms on replicated counters a template may have in this kernel */

struct __align__(16) MultiShared {
  uint32_t wtop[LEAN_WARPS][MULTI_M];               // per-warp top-M (compact) keys of this wave
  int32_t gt_c1[MULTI_GT][4];                       // (16-byte aligned) ... and {limit, payload shift, payload mask, counter base}
  int32_t gt_commit[MULTI_GT][4];                   // (16-byte aligned) per replicated-counter term: {counter base, inc, PTS constraint tracked or -1, n_present}
  uint32_t rmax[LEAN_WARPS], rbar[LEAN_WARPS];      // replay: per-warp maxima of a round
  int32_t wfeas[LEAN_WARPS];
  int32_t gt_term[MULTI_GT];                        // indices of the Filter terms that read a replicated (non node-local) counter
  uint32_t gt_shift[MULTI_GT], gt_mask[MULTI_GT];   // ... and where their domain id sits in the payload
  int32_t acc_node[MULTI_MAX_ACC];                  // replay: nodes accepted in this wave, in order
  int32_t full[MULTI_GT];                           // replay: counter cell of term q that the last commit pushed over its limit, or -2
  int32_t n_gt, accepted, dead, stopb;
  int32_t single_use, pad_ms[3];
};

__shared__ MultiShared ms;

struct MultiParams {
  uint32_t pay_shift[LEAN_MAX_SLOTS];   // record slot s (a topology column) -> bit position of its dom+1 field in the payload
  uint32_t pay_mask[LEAN_MAX_SLOTS];    // field mask (0: the slot is a node-local counter, not carried)
};

// 32-bit keys for everything inside this kernel: (score+1) in bits 20..31, (2^20-1 - global index) below (needs N < 2^20 and
// score+1 < 4096, both host-checked). Same order as pack_key (highest score, then lowest index); one REDUX per arg-max.
#define MULTI_IDX_BITS 20
#define MULTI_IDX_MASK ((1u << MULTI_IDX_BITS) - 1u)
__device__ __forceinline__ uint32_t ckey(int32_t score, uint32_t gidx) { return ((uint32_t)(score + 1) << MULTI_IDX_BITS) | (MULTI_IDX_MASK - gidx); }
__device__ __forceinline__ int32_t ckey_index(uint32_t ck) { return (int32_t)(MULTI_IDX_MASK - (ck & MULTI_IDX_MASK)); }

// Shared-memory accesses of the replay loop by explicit 32-bit shared address: nvcc otherwise re-derives the CTA's shared
// window base (S2UR SR_CgaCtaId + LEA) in front of every access inside these barrier-separated regions.
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
#define MS_OFF(field) ((uint32_t)offsetof(MultiShared, field))

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
  unsigned long long *c_pay = reinterpret_cast<unsigned long long *>(c_npods + cp);   // (chunk_pad is a multiple of 4: 8-byte aligned)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const int32_t lo = min(p.n, cta * p.chunk), hi = min(p.n, lo + p.chunk);
  const int32_t cnt_nodes = hi - lo;      // <= LEAN_THREADS (host-checked): one node per thread
  // shared address of `ms`, computed once; the volatile move keeps nvcc from re-deriving it (S2UR SR_CgaCtaId) at every use
  uint32_t msb;
  { const uint32_t t0 = (uint32_t)__cvta_generic_to_shared(&ms); asm volatile("mov.u32 %0, %1;" : "=r"(msb) : "r"(t0)); }
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
    // the node's payload for the candidate exchange: domain id + 1 of every topology slot (static: labels do not change)
    unsigned long long py = 0ull;
    for (int s = 0; s < lp.n_slots; s++)
      if (mp.pay_mask[s]) py |= (unsigned long long)((uint32_t)(r4[10 + s] + 1) & mp.pay_mask[s]) << mp.pay_shift[s];
    c_pay[j] = py;
  }
  for (int k = tid; k < (int)(sizeof(ccsim_template) / 8); k += LEAN_THREADS)
    reinterpret_cast<unsigned long long *>(&ls.tmpl)[k] = reinterpret_cast<const unsigned long long *>(&p.templates[0])[k];
  for (int j = 0; j < p.n_counters; j++) {
    const DevCounter &dc = p.counters[j];
    if (dc.topo_col < 0) continue;
    for (int d = tid; d < dc.n_domains; d += LEAN_THREADS) smem_cnt[dc.smem_off + d] = dc.init[d];
  }
  if (tid == 0) { ls.aff_total = p.templates[0].aff_total_init; ls.winner = -1; ls.stop = 0; ls.dirty = 1; ms.accepted = 0; ms.dead = 0; ms.stopb = 0; ms.n_gt = 0; }
  __syncthreads();
  for (int c = 0; c < ls.tmpl.n_pts; c++) lean_pts_recount(p, smem_cnt, c);

#ifdef CCSIM_PHASE_TIMERS
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc0 = 0, tc1 = 0;
#endif
  long long k = 0, wv = 0;
  bool limit_hit = false;   // postBindHook's limit (simulator.go:300-305)
  uint32_t wtag = 1;
  uint32_t tag = (p.epoch << 12) | wtag;
  for (;; wv++) {
    PH_START();
    if (p.max_pods > 0 && k >= p.max_pods) { limit_hit = true; break; }   // uniform; no shared write (slower threads may still be reading ls.stop)
    if (k > p.pod_cap) { if (tid == 0) ls.stop = 3; __syncthreads(); break; }
    if (ls.dirty) {
      if (tid == 0) {
        lean_build_consts(p, lp);
        int g = 0;
        for (int q = 0; q < ls.n_cmp_terms; q++)
          if (ls.terms[q].cnt_off >= 0 && g < MULTI_GT) {
            const int sl = ls.terms[q].slot - 10;
            ms.gt_shift[g] = mp.pay_shift[sl]; ms.gt_mask[g] = mp.pay_mask[sl];
            ms.gt_c1[g][0] = ls.terms[q].lim; ms.gt_c1[g][1] = (int32_t)mp.pay_shift[sl]; ms.gt_c1[g][2] = (int32_t)mp.pay_mask[sl]; ms.gt_c1[g][3] = ls.terms[q].cnt_off;
            ms.gt_commit[g][0] = ls.terms[q].cnt_off; ms.gt_commit[g][1] = 0; ms.gt_commit[g][2] = -1; ms.gt_commit[g][3] = 0;
            for (int j = 0; j < p.n_counters; j++)
              if (p.counters[j].topo_col >= 0 && p.counters[j].smem_off == ls.terms[q].cnt_off) {
                ms.gt_commit[g][1] = ls.cinfo[j].inc; ms.gt_commit[g][2] = ls.cinfo[j].pts_idx; ms.gt_commit[g][3] = ls.cinfo[j].n_present;
              }
            ms.gt_term[g++] = q;
          }
        ms.n_gt = g;
        int su1 = 0;   // a self-matching required anti-affinity term on a node-local counter: count 0 -> inc > limit 0 after one clone
        for (int q = 0; q < ls.n_cmp_terms; q++)
          if (ls.terms[q].cnt_off < 0 && ls.terms[q].kind == LT_ANTI)
            for (int j = 0; j < p.n_counters; j++)
              if (p.counters[j].topo_col < 0 && 10 + lp.counter_slot[j] == ls.terms[q].slot && ls.cinfo[j].inc > 0) su1 = 1;
        ms.single_use = su1;
      }
      __syncthreads();
      if (tid == 0) ls.dirty = 0;
    }
    // ---- fused Filter pass: this thread's node ----
    uint32_t key = 0u;
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
        key = ckey(sc, (uint32_t)(p.node_base + lo + j));
      }
    }
    // ---- the warp's M best keys (REDUX rounds; keys are unique, 0 = none) ----
    {
      uint32_t rem = key;
      const int nf = __popc(__ballot_sync(0xffffffffu, key != 0u));
      if (lane == 0) ms.wfeas[warp] = nf;
      #pragma unroll
      for (int r = 0; r < MULTI_M; r++) {
        uint32_t v = 0u;
        if (r < nf) {                         // warp-uniform: no REDUX rounds for entries that do not exist
          v = __reduce_max_sync(0xffffffffu, rem);
          if (rem == v) rem = 0u;
        }
        if (lane == 0) ms.wtop[warp][r] = v;
      }
    }
    PH_MARK(0);
    __syncthreads();                                                    // S1
    PH_MARK(1);
    const unsigned long long tagbits = (unsigned long long)tag << KEY_TAG_SHIFT;
    if (warp == 0) {
      // ---- the CTA's M best: merge of the 24 sorted warp lists (lane w walks warp w's list), then publish the pairs ----
      int32_t total = lane < LEAN_WARPS ? ms.wfeas[lane] : 0;
      total = __reduce_add_sync(0xffffffffu, total);
      int ptr = 0, L = 0;
      uint32_t mykey = 0u;
      uint32_t head = lane < LEAN_WARPS ? ms.wtop[lane][0] : 0u;
      for (int r = 0; r < MULTI_M; r++) {
        const uint32_t g = __reduce_max_sync(0xffffffffu, head);
        if (g == 0u) break;                    // uniform
        if (head == g) { ptr++; head = ptr < MULTI_M ? ms.wtop[lane][ptr] : 0u; }
        if (lane == r) mykey = g;
        L = r + 1;
      }
      unsigned long long pay = 0ull;
      if (lane < L) {
        const int32_t jj = ckey_index(mykey) - (p.node_base + lo);
        pay = c_pay[jj];
        // The node's key after one more clone ("second life" in the replay): the node-local part of the Filter pass and the
        // score again, on the row as it would be after this commit (types.go:409-427). Per-domain terms are re-checked by
        // the replay itself. 0 = the node would not take another clone.
        if (!ms.single_use) {    // (a clone that blocks its own node — hostname anti-affinity — never has a second life)
        const uint4 *r = rec + (size_t)jj * su;
        const uint4 u1 = r[1], u2 = r[2];
        const long long free_cpu = (long long)(((unsigned long long)u1.y << 32) | u1.x) - ls.tmpl.req_cpu;
        const long long free_mem = (long long)(((unsigned long long)u1.w << 32) | u1.z) - ls.tmpl.req_mem;
        bool ok2 = (free_cpu >= ls.eq_cpu) & (free_mem >= ls.eq_mem) & ((int32_t)u2.x - 1 >= ls.pods_need);
        const int32_t *r4 = reinterpret_cast<const int32_t *>(r);
        for (int q = 0; q < ls.n_cmp_terms; q++) {
          const LeanTerm lt = ls.terms[q];
          if (lt.cnt_off >= 0) continue;                       // replicated counters: the replay's business
          int inc = 0;
          for (int j = 0; j < p.n_counters; j++) if (p.counters[j].topo_col < 0 && 10 + lp.counter_slot[j] == lt.slot) inc = ls.cinfo[j].inc;
          ok2 &= (r4[lt.slot] + inc <= lt.lim);
        }
        if (ok2) {
          const int32_t sc2 = score_node(c_acpu[jj], c_amem[jj], c_zcpu[jj] + ls.tmpl.nz_cpu + ls.tmpl.least_cpu, c_zmem[jj] + ls.tmpl.nz_mem + ls.tmpl.least_mem,
                                         c_rcpu[jj] + ls.tmpl.req_cpu + ls.tmpl.bal_cpu, c_rmem[jj] + ls.tmpl.req_mem + ls.tmpl.bal_mem, ls.sw);
          pay |= (unsigned long long)(uint32_t)(sc2 + 1) << MULTI_NEXT_SHIFT;
        }
        }
      }
      pay |= ((unsigned long long)L << MULTI_LEN_SHIFT) | ((unsigned long long)(total > L ? 1 : 0) << MULTI_MORE_BIT);
      unsigned long long *myslots = p.slots + ((size_t)(wv & 1) * CCSIM_MAX_GRID + cta) * SLOT_STRIDE;
      if (lane < MULTI_M) {
        st_slot(&myslots[2 * lane], (unsigned long long)mykey | tagbits);
        st_slot(&myslots[2 * lane + 1], pay | tagbits);
      }
    }
    PH_MARK(2);
    // ---- gather: every (CTA, entry) pair is polled by one thread and stays in its registers for the replay ----
    const int e0 = tid, e1 = tid + LEAN_THREADS;
    unsigned long long a0w = 0ull, b0 = 0ull, a1w = 0ull, b1 = 0ull;
    {
      const unsigned long long *base = p.slots + (size_t)(wv & 1) * CCSIM_MAX_GRID * SLOT_STRIDE;
      const int tot = p.grid * MULTI_M;
      // (1) one lane per CTA line waits for that line's first key word — 148 pollers per line grid-wide, as in the lean kernel;
      //     letting all 768 threads spin on their own entries (1184 pollers per line) delays the very stores they wait for
      if (warp < (p.grid + 31) / 32) {
        const int c = warp * 32 + lane;
        unsigned spins = 0;
        bool pending;
        do {
          pending = (c < p.grid) && ((uint32_t)(ld_slot(base + (size_t)c * SLOT_STRIDE) >> KEY_TAG_SHIFT) != tag);
          if (++spins > WATCHDOG_SPINS) { ms.dead = 1; break; }
        } while (__any_sync(0xffffffffu, pending));
      }
      __syncthreads();
      // (2) every thread fetches its entries; words of a line are separate stores, so each is still validated by its own tag
      bool need0 = e0 < tot, need1 = e1 < tot;
      unsigned spins = 0;
      while ((need0 | need1) && !ms.dead) {
        if (need0) ld_slot2(base + (size_t)(e0 >> 3) * SLOT_STRIDE + 2 * (e0 & 7), a0w, b0);
        if (need1) ld_slot2(base + (size_t)(e1 >> 3) * SLOT_STRIDE + 2 * (e1 & 7), a1w, b1);
        if (need0 && (uint32_t)(a0w >> KEY_TAG_SHIFT) == tag && (uint32_t)(b0 >> KEY_TAG_SHIFT) == tag) need0 = false;
        if (need1 && (uint32_t)(a1w >> KEY_TAG_SHIFT) == tag && (uint32_t)(b1 >> KEY_TAG_SHIFT) == tag) need1 = false;
        if (++spins > WATCHDOG_SPINS) { ms.dead = 1; break; }
      }
      if (need0 | need1) { a0w = b0 = a1w = b1 = 0ull; }
      b0 &= KEY_BODY_MASK; b1 &= KEY_BODY_MASK;
      if (e0 >= tot) { a0w = 0ull; b0 = 0ull; }
      if (e1 >= tot) { a1w = 0ull; b1 = 0ull; }
    }
    const uint32_t a0 = (uint32_t)a0w, a1 = (uint32_t)a1w;      // compact keys as published (the tag sits above bit 44)
    uint32_t k0 = a0, k1 = a1;                                   // ... and as they stand in the replay (a winner comes back once with its next key)
    bool second0 = false, second1 = false;
    // counter cells each of my candidates depends on (-1: the node lacks the key and the term lets it pass)
    int32_t i0[MULTI_GT], i1[MULTI_GT];
    const int n_gt = ms.n_gt;
    const bool has1 = (e1 & ~31) < p.grid * MULTI_M;
    #pragma unroll
    for (int q = 0; q < MULTI_GT; q++) {
      i0[q] = -1; i1[q] = -1;
      if (q < n_gt) {
        const int4 c1 = *reinterpret_cast<const int4 *>(&ms.gt_c1[q][0]);      // {lim, shift, mask, counter base}
        const int32_t v0 = (int32_t)((uint32_t)(b0 >> c1.y) & (uint32_t)c1.z) - 1;
        i0[q] = v0 >= 0 ? c1.w + v0 : -1;
        if (has1) {          // warp-uniform: only the first grid*8 - 768 threads hold a second candidate
          const int32_t v1 = (int32_t)((uint32_t)(b1 >> c1.y) & (uint32_t)c1.z) - 1;
          i1[q] = v1 >= 0 ? c1.w + v1 : -1;
        }
      }
    }
    PH_MARK(3);
    // ---- replay: the reference cycles k, k+1, ... this wave can decide; every CTA does the same, all threads take part ----
    // Each round: every thread re-checks its (<= 2) candidates against the CTA's counter replicas (a candidate that fails is
    // dropped for the rest of the wave: monotone), block arg-max over the survivors, the thread holding the maximum commits.
    bool live0 = a0 != 0u, live1 = a1 != 0u;
    uint32_t bar = 0u;      // highest "last published key" of a list that ran dry while its CTA has more nodes
    int32_t acc = 0;
    const int L0 = (int)((b0 >> MULTI_LEN_SHIFT) & 15ull), L1 = (int)((b1 >> MULTI_LEN_SHIFT) & 15ull);
    const bool more0 = (b0 >> MULTI_MORE_BIT) & 1ull, more1 = (b1 >> MULTI_MORE_BIT) & 1ull;
    const unsigned grp = 0xffu << (lane & ~7);
    __syncthreads();                                                    // S2: ms.dead, the counters of the previous wave's recount
    const bool dead = ms.dead != 0;
    // Round 0 needs no check: every published candidate passed the scan under the very counts the replicas hold now. After a
    // commit only the candidates sitting in a counter cell that has just gone over its limit die; the committing warp
    // names those cells (ms.full), everybody else compares. A warp whose candidates did not change keeps its maxima.
    uint32_t wm = 0u, wb = 0u;
    bool refresh = true;                 // warp-uniform: recompute this warp's maxima
    for (int round = 0; !dead; round++) {
      if (round > 0) {
        bool died = false;
        #pragma unroll
        for (int q = 0; q < MULTI_GT; q++) {
          if (q < n_gt) {
            const int32_t f = (int32_t)lds_u32(msb + MS_OFF(full) + 4u * q);
            if (f >= 0) {
              if (live0 && i0[q] == f) { live0 = false; died = true; }
              if (has1 && live1 && i1[q] == f) { live1 = false; died = true; }
            }
          }
        }
        refresh |= __any_sync(0xffffffffu, died);
      }
      if (refresh) {
        const uint32_t v0 = live0 ? k0 : 0u, v1 = live1 ? k1 : 0u;
        // a list (8 consecutive lanes) without a live entry whose CTA has unpublished feasible nodes: those rank below its last key
        // (only entries in their first life count: a CTA's unpublished nodes rank below its last PUBLISHED key, and a winner that
        //  came back with its next key may well rank below that)
        const unsigned bal0 = __ballot_sync(0xffffffffu, live0 && !second0), bal1 = __ballot_sync(0xffffffffu, live1 && !second1);
        if (!(bal0 & grp) && more0 && (e0 & 7) == L0 - 1) bar = a0 > bar ? a0 : bar;
        if (!(bal1 & grp) && more1 && (e1 & 7) == L1 - 1) bar = a1 > bar ? a1 : bar;
        wm = __reduce_max_sync(0xffffffffu, v0 > v1 ? v0 : v1);
        wb = __reduce_max_sync(0xffffffffu, bar);
        if (lane == 0) { sts_u32(msb + MS_OFF(rmax) + 4u * warp, wm); sts_u32(msb + MS_OFF(rbar) + 4u * warp, wb); }
        refresh = false;
      }
      __syncthreads();                                                  // A
#ifdef CCSIM_PHASE_TIMERS
      if (round == 0) { PH_MARK(1); } else { PH_MARK(6); }
#endif
      const uint32_t g = __reduce_max_sync(0xffffffffu, lane < LEAN_WARPS ? lds_u32(msb + MS_OFF(rmax) + 4u * lane) : 0u);
      const uint32_t gb = __reduce_max_sync(0xffffffffu, lane < LEAN_WARPS ? lds_u32(msb + MS_OFF(rbar) + 4u * lane) : 0u);
      if (g == 0u || g < gb) break;            // block-uniform: nothing left, or an unpublished node could rank above g
      if (wm == g) {      // warp-uniform: this warp holds the winner; it commits like the lean kernel's warp 0 does
        // ---- commit pod k+acc (assume -> AssumePod -> NodeInfo.update(+1): schedule_one.go:967-984, types.go:409-427) ----
        const bool own0 = live0 && k0 == g, own1 = live1 && k1 == g;
        const int ol = __ffs(__ballot_sync(0xffffffffu, own0 | own1)) - 1;
        const unsigned long long pay = __shfl_sync(0xffffffffu, own0 ? b0 : b1, ol);
        // the winner comes back once with the key it has after this clone (if it still fits); when a node wins for the second
        // time in a wave its third key is unknown: the wave ends after that commit
        const bool was_second = __any_sync(0xffffffffu, (own0 && second0) || (own1 && second1));
        if (own0) {
          const uint32_t ns = (uint32_t)(b0 >> MULTI_NEXT_SHIFT) & 0xfffu;
          if (ns && !second0) { k0 = (ns << MULTI_IDX_BITS) | (k0 & MULTI_IDX_MASK); second0 = true; } else live0 = false;
        }
        if (own1) {
          const uint32_t ns = (uint32_t)(b1 >> MULTI_NEXT_SHIFT) & 0xfffu;
          if (ns && !second1) { k1 = (ns << MULTI_IDX_BITS) | (k1 & MULTI_IDX_MASK); second1 = true; } else live1 = false;
        }
        // Only what the next round depends on happens here: the counter cells of the winner's domains (lane q = term q; the
        // host guarantees one term per incremented replicated counter), whether a cell went over its limit, whether a PTS
        // minimum moved. The winner's row (NodeInfo.update, node-local counters) is brought up to date after the last round.
        bool minchg = false;
        if (lane < n_gt) {
          const int4 gc = *reinterpret_cast<const int4 *>(&ms.gt_commit[lane][0]);   // {cnt_off, inc, pts_idx, n_present}
          const int4 c1 = *reinterpret_cast<const int4 *>(&ms.gt_c1[lane][0]);       // {lim, shift, mask, cnt_off}
          const int32_t lim = c1.x;
          const int32_t v = (int32_t)((uint32_t)(pay >> c1.y) & (uint32_t)c1.z) - 1;
          int32_t fullcell = -2;
          if (v >= 0) {
            const int32_t old = smem_cnt[gc.x + v], nv = old + gc.y;
            smem_cnt[gc.x + v] = nv;
            if (nv > lim) fullcell = gc.x + v;             // candidates in this cell are dead from now on
            if (gc.y && gc.z >= 0 && v < gc.w && old == ls.ptsmin[gc.z]) {
              const int32_t left = ls.ptsnum[gc.z] - 1;
              ls.ptsnum[gc.z] = left;
              minchg = left <= 0;                          // the global minimum of this constraint moves: limits change, rescan
            }
          }
          ms.full[lane] = fullcell;
        }
        bool stopb = __any_sync(0xffffffffu, minchg);
        stopb |= (p.max_pods > 0 && k + acc + 1 >= p.max_pods);
        stopb |= (k + acc + 1 >= p.pod_cap) | (acc + 1 >= MULTI_MAX_ACC) | was_second;
        if (lane == 0) { ms.stopb = stopb ? 1 : 0; ms.acc_node[acc] = ckey_index(g); }
        refresh = true;
      }
      acc++;
      __syncthreads();                                                  // B: counters, ms.stopb
      PH_MARK(7);
      if (lds_u32(msb + MS_OFF(stopb))) break;
    }
    __syncthreads();     // every thread has left the loop (some after barrier A, some after B)
    // ClusterCapacityBinder.Bind + postBindHook: pod k+i -> node (plugin.go:34-53; simulator.go:297-312). Every CTA knows the
    // whole list; CTA 0 records it.
    if (cta == 0 && tid < acc && k + tid < p.pod_cap) p.pod_node[k + tid] = ms.acc_node[tid];
    // ---- assume -> AssumePod -> NodeInfo.update(+1) (schedule_one.go:967-984, types.go:409-427) for the winners this CTA owns:
    //      one thread per accepted pod (a node is accepted at most once per wave) ----
    if (tid >= 32 && tid - 32 < acc) {       // warp 1 (and up): warp 0 recounts PTS minima meanwhile
      const ccsim_template &t = ls.tmpl;
      const int i = tid - 32;
      const int32_t w = ms.acc_node[i] - p.node_base;
      // a node is accepted at most twice per wave, the second time as the wave's last commit: its first thread applies both
      const bool twice = (i != acc - 1) && ms.acc_node[acc - 1] == ms.acc_node[i];
      const bool skip = (i == acc - 1) && [&] { for (int q = 0; q < acc - 1; q++) if (ms.acc_node[q] == ms.acc_node[i]) return true; return false; }();
      const int mult = twice ? 2 : 1;
      if (w >= lo && w < hi && !skip) {
        const int32_t jw = w - lo;
        const long long rc = c_rcpu[jw] + mult * t.req_cpu, rm = c_rmem[jw] + mult * t.req_mem;
        const int32_t np = c_npods[jw] + mult;
        c_rcpu[jw] = rc; c_rmem[jw] = rm; c_zcpu[jw] += mult * t.nz_cpu; c_zmem[jw] += mult * t.nz_mem; c_npods[jw] = np;
        unsigned long long *r8 = reinterpret_cast<unsigned long long *>(rec + (size_t)jw * su);
        int32_t *r4 = reinterpret_cast<int32_t *>(r8);
        r8[2] = (unsigned long long)(c_acpu[jw] - rc);
        r8[3] = (unsigned long long)(c_amem[jw] - rm);
        r4[8] = c_apods[jw] - np;
        r4[9] = -1;            // this node's NodeInfo generation changed: its memoised score is stale
        for (int j = 0; j < p.n_counters; j++) {
          const CommitInfo ci = ls.cinfo[j];
          if (ci.inc && ci.local) r4[10 + lp.counter_slot[j]] += mult * ci.inc;   // node-local counters (written back when the run ends)
        }
      }
    }
    if (warp == 0) {
      const ccsim_template &t = ls.tmpl;
      // a PTS minimum moved: recount it here and move the term's limit, instead of a block-wide recount + rebuild of all
      // Filter constants (filtering.go:56-69,98-137: minMatchNum / criticalPaths)
      for (int c = 0; c < t.n_pts; c++) {
        if (t.pts[c].min_zero || ls.ptsnum[c] > 0) continue;
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
            for (int q = 0; q < ms.n_gt; q++) if (ms.gt_term[q] == c) ms.gt_c1[q][0] = ls.terms[c].lim;
          }
        }
        __syncwarp();
      }
      if (lane == 0) {
        ms.accepted = acc;
        ms.stopb = 0;
        if (dead) ls.stop = 3;
        else if (acc == 0) ls.stop = 1;          // no feasible node anywhere: the pod is unschedulable
      }
    }
    PH_MARK(4);
    __syncthreads();                                                    // S3
    PH_MARK(5);
    k += ms.accepted;
    if (ls.stop) break;
    for (int c = 0; c < ls.tmpl.n_pts; c++)
      if (!ls.tmpl.pts[c].min_zero && ls.ptsnum[c] <= 0 && p.counters[ls.tmpl.pts[c].counter].n_present > 0) lean_pts_recount(p, smem_cnt, c);
    wtag = (wtag == 4095u) ? 1u : wtag + 1u;
    tag = (p.epoch << 12) | wtag;
  }

  // ---- write the tile back: the global columns are the snapshot-after-run (terminal diagnosis, ccsim_node_counts) ----
  for (int32_t j = tid; j < cnt_nodes; j += LEAN_THREADS) {
    const int32_t i = lo + j;
    p.req_cpu[i] = c_rcpu[j]; p.req_mem[i] = c_rmem[j]; p.nz_cpu[i] = c_zcpu[j]; p.nz_mem[i] = c_zmem[j]; p.npods[i] = c_npods[j];
    const int32_t *r4 = reinterpret_cast<const int32_t *>(rec + (size_t)j * su);
    for (int sl = 0; sl < lp.n_slots; sl++) if (lp.slot_topo[sl] < 0) p.counters[lp.slot_counter[sl]].work[i] = r4[10 + sl];
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
      o->stop_code = limit_hit ? CCSIM_STOP_LIMIT_REACHED : CCSIM_STOP_UNSCHEDULABLE;
      o->error = (ls.stop == 3) ? 1 : 0;
      o->waves = limit_hit ? wv : wv + 1;
      o->evals = o->waves * (long long)p.n;
      o->examined = o->evals;
      for (int c = 0; c < CCSIM_MAX_PTS; c++) o->ptsmin[c] = ls.ptsmin[c];
      o->aff_total = ls.aff_total;
#ifdef CCSIM_PHASE_TIMERS
      for (int q = 0; q < 8; q++) o->phase_cycles[q] = ph[q];
#endif
    }
  }
}
