// ccsim_multi.cuh — multi-commit waves for templates coupled through per-domain counters (PodTopologySpread DoNotSchedule,
// required pod anti-affinity): several reference scheduling cycles per grid-wide exchange, identical pod -> node sequence,
// on one GPU or over node shards on several GPUs (the exchange then crosses NVLink inside the same kernel).
//
// Why it is legal. Within a stretch of cycles in which no PodTopologySpread global minimum changes, a node's feasibility
// depends on its own row and on the match counts of its topology domains only, and it is MONOTONE: counts only grow
// (inc >= 0), so a feasible node can become infeasible but never the reverse; a node's score changes only when the node
// itself is committed. The winner of each cycle is therefore the highest-keyed node that is still feasible under the
// counts as updated so far (podtopologyspread/filtering.go:311-356, interpodaffinity/filtering.go:352-432;
// schedule_one.go:894-941 with the canonical first-max tie rule). Every CTA publishes its M best feasible nodes (the exact
// top-M of its tile: keys are unique) with the domain ids their feasibility depends on; every CTA of every rank then
// replays those cycles redundantly and deterministically from the same published data.
//
// Which published candidates may be replayed. A CTA with more than M feasible nodes has UNSEEN nodes, all keyed below its
// M-th (last) published key. T = the largest such last key over all lists is a wave-wide bar: a candidate keyed >= T
// outranks every unseen node of every tile, a candidate keyed below T might not. The replay therefore works on the
// candidates keyed >= T only (compacted into shared memory) and ends when the best live key falls below T, when
//   (a) a PTS minimum moves (limits change, rejected nodes may come back: rescan),
//   (b) the pod limit is reached (simulator.go:300-305),
//   (c) a node wins for the second time (see below).
// A committed node may win again inside the wave with its new score, which is in nobody's list (a randomized differential
// test, tests/test_gpu_stress.py, found this hole in the first version: spread-only templates). Every published candidate
// therefore carries, next to its domain ids, its key after one more clone (node-local Filter part + score recomputed by the
// publisher; 0 = it would not fit). The winner comes back into the replay once with that key ("second life"); when a node
// wins for the second time in a wave its third key is unknown and the wave ends after that commit.
// The best candidate of a wave is always >= T and feasible, so every wave makes progress; a wave without candidates is the
// Unschedulable stop. Node-local terms (hostname anti-affinity, ...) need no re-check: a node appears once per wave.
//
// Exchange: one 128-byte line per CTA = M (key, payload) pairs, each word tagged (self-validating, no fences).
//   single GPU : lines live in the handle's slot buffer (L2), st/ld.relaxed.gpu;
//   node shards: every CTA stores its line into EVERY rank's line buffer (CUDA IPC mappings, st.relaxed.sys.v2 over NVLink)
//                and polls its local copy — an all-gather of world x grid lines fused into the persistent kernel.
// Replay: ONE warp, no barrier inside: <= 8 candidates per lane in registers; a round is arg-max (REDUX) -> commit (lane q =
//   counter term q: cell += inc, over-limit test, PTS-minimum tracking, all from registers) -> kill the candidates sitting in
//   a cell that just filled (field compare against the shuffled cell id). ~250 cycles per reference cycle instead of the
//   ~2000 of a block-wide round (two barriers over 24 warps); the other 23 warps wait at the barrier that ends the wave.
//   Row updates of the winners are done by each node's own thread after that barrier (a thread owns its node).
// Look-ahead waves: a PodTopologySpread minimum move that REOPENS closed domains would end the wave (the reopened nodes were
//   rejected by the scan and are in nobody's list). When a term's limit is about to move, the scan publishes the nodes of its
//   closed cells too; they sit in the replay as dormant candidates (key 0) and are rebuilt from the wave's candidate arrays when
//   the move comes — see "dormant candidates" in the replay. C4: 3654 -> 2260 waves (scripts/wave_sim.py models the wave structure
//   on the CPU and was used to choose the rule).
#pragma once
#include "ccsim_lean.cuh"

#define MULTI_M 8                 /* candidates per CTA and wave: 8 x 16 B = the CTA's 128-byte slot line */
#define MULTI_PAY_BITS 27         /* payload bits for domain ids (dom+1 per topology slot) */
#define MULTI_NEXT_SHIFT 27       /* 12 bits: (score + 1) of the node after one more clone, 0 = it would not fit any more */
#define MULTI_MORE_BIT 32         /* key word of the last entry: the tile has more feasible nodes than it published */
#define MULTI_MAX_ACC 64          /* commits one wave may decide */
#define MULTI_GT 6                /* Filter terms on replicated counters a template may have in this kernel */
#ifndef MULTI_CPT
#define MULTI_CPT 4               /* candidates per replay lane */
#endif
#define MULTI_CAP (32 * MULTI_CPT)
#define MULTI_BINS 64
#define MULTI_RELAX_K 8           /* look-ahead on a PTS term when at most this many of its domains still sit at the global minimum ... */
#define MULTI_RELAX_R 3           /* ... nodes in cells up to this far over the limit are published as dormant candidates */
#define MULTI_XPT (((CCSIM_MAX_WORLD - 1) * MULTI_CAP + LEAN_THREADS - 1) / LEAN_THREADS)   /* node shards: remote candidates per thread */

// cross-GPU line buffers inside every rank's exchange allocation (64-bit words): [parity][source rank][CTA][16]
#define XLEAN_WORDS (2 * CCSIM_MAX_WORLD * SLOT_STRIDE)
#define XLINES_OFF XLEAN_WORDS
#define XLINES_WORDS (2 * CCSIM_MAX_WORLD * CCSIM_MAX_GRID * SLOT_STRIDE)
#define XSLOTS_TOTAL_WORDS (XLEAN_WORDS + XLINES_WORDS)

struct __align__(16) MultiShared {
  uint32_t wtop[LEAN_WARPS][MULTI_M];               // per-warp top-M (compact) keys of this wave
  int32_t gt_c1[MULTI_GT][4];                       // per replicated-counter term: {limit, payload shift, payload mask, domains of the counter}
  int32_t gt_commit[MULTI_GT][4];                   // ... {counter base, inc, PTS constraint tracked or -1, n_present}
  uint32_t ckey[MULTI_CAP], cdom[MULTI_CAP], cnext[MULTI_CAP];   // the wave's candidates keyed >= T (unordered)
  uint32_t red[LEAN_WARPS], red2[LEAN_WARPS];       // block reductions (T, best key)
  uint32_t hist[MULTI_BINS];
  int32_t wfeas[LEAN_WARPS];
  int32_t gt_term[MULTI_GT];                        // indices of the Filter terms that read a replicated (non node-local) counter
  int32_t acc_node[MULTI_MAX_ACC];                  // replay: nodes accepted in this wave, in order
  int32_t n_gt, accepted, dead, stopb;
  int32_t single_use, ncand;
  int32_t xcount[CCSIM_MAX_WORLD];                  // node shards: candidates in each rank's summary
  uint32_t xglob[4];                                // node shards: best key, T_list, bar over all ranks
  uint32_t delta, pad_ms;
  int32_t relax[LEAN_MAX_TERMS];                    // per Filter term: this wave's look-ahead over the limit (0: strict), see "dormant candidates"
  int32_t force_strict, st_relaxed, st_empty, pad_r;
  long long ph[8], tc0, st_cand, st_overflow, st_rounds;       // CTA 0 / thread 0: clock cycles per phase, replay statistics
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

template <bool XGPU> __device__ __forceinline__ void ld_line2(const unsigned long long *p, unsigned long long &a, unsigned long long &b) {
  if (XGPU) asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_line2_sys(unsigned long long *p, unsigned long long a, unsigned long long b) {
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// position of the g-th (0-based, g < 4) set bit of m, or -1
__device__ __forceinline__ int nth_set_lane(unsigned m, int g) {
  #pragma unroll
  for (int i = 0; i < 3; i++) if (i < g) m &= m - 1;
  return m ? __ffs(m) - 1 : -1;
}

// warp-aggregated append of the lanes' candidates to the wave's candidate arrays (order does not matter: keys are unique)
__device__ __forceinline__ void multi_append(bool keep, uint32_t ck, unsigned long long b, int lane) {
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m) {
    int base = 0;
    const int leader = __ffs(m) - 1;
    if (lane == leader) base = atomicAdd(&ms.ncand, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    const int idx = base + __popc(m & ((1u << lane) - 1u));
    if (keep && idx < MULTI_CAP) {
      ms.ckey[idx] = ck;
      ms.cdom[idx] = (uint32_t)b & ((1u << MULTI_PAY_BITS) - 1u);
      ms.cnext[idx] = (uint32_t)(b >> MULTI_NEXT_SHIFT) & 0xfffu;
    }
  }
}

// phase timers live in shared memory (thread 0 of CTA 0 only): registers are what this kernel is short of
#define MPH_START() do { if (cta == 0 && tid == 0) ms.tc0 = clock64(); } while (0)
#define MPH_MARK(i) do { if (cta == 0 && tid == 0) { const long long tc1_ = clock64(); ms.ph[i] += tc1_ - ms.tc0; ms.tc0 = tc1_; } } while (0)

template <bool XGPU>
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
  const int su = lp.stride_u;
  const uint32_t cnt_sa = pin_u32(smem_u32(smem_cnt)), ms_sa = pin_u32(smem_u32(&ms));   // shared bases for the replay's explicit-address accesses
#define MS_SA(field) (ms_sa + (uint32_t)offsetof(MultiShared, field))
  const int nlists = p.grid;                            // lines of this GPU (node shards: every rank launches the same grid, sized from the largest shard)
  const int tot = nlists * MULTI_M;

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
  if (tid == 0) { ls.aff_total = p.templates[0].aff_total_init; ls.winner = -1; ls.stop = 0; ls.dirty = 1; ms.accepted = 0; ms.dead = 0; ms.stopb = 0; ms.n_gt = 0; ms.ncand = 0; ms.delta = 1u << MULTI_IDX_BITS; ms.force_strict = 0; ms.st_relaxed = 0; ms.st_empty = 0;
                  for (int q = 0; q < LEAN_MAX_TERMS; q++) ms.relax[q] = 0;
                  for (int q = 0; q < 8; q++) ms.ph[q] = 0; ms.tc0 = 0; ms.st_cand = 0; ms.st_overflow = 0; ms.st_rounds = 0; }
  __syncthreads();
  for (int c = 0; c < ls.tmpl.n_pts; c++) lean_pts_recount(p, smem_cnt, c);

  long long k = 0, wv = 0;
  uint32_t delta = 1u << MULTI_IDX_BITS;     // bar distance below the best key: starts at one score level
  bool limit_hit = false;   // postBindHook's limit (simulator.go:300-305)
  uint32_t wtag = 1;
  uint32_t tag = (p.epoch << 12) | wtag;
  for (;; wv++) {
    MPH_START();
    if (p.max_pods > 0 && k >= p.max_pods) { limit_hit = true; break; }   // uniform; no shared write (slower threads may still be reading ls.stop)
    if (k > p.pod_cap) { if (tid == 0) ls.stop = 3; __syncthreads(); break; }
    if (ls.dirty) {
      if (tid == 0) {
        lean_build_consts(p, lp);
        int g = 0;
        for (int q = 0; q < ls.n_cmp_terms; q++)
          if (ls.terms[q].cnt_off >= 0 && g < MULTI_GT) {
            const int sl = ls.terms[q].slot - 10;
            ms.gt_c1[g][0] = ls.terms[q].lim; ms.gt_c1[g][1] = (int32_t)mp.pay_shift[sl]; ms.gt_c1[g][2] = (int32_t)mp.pay_mask[sl]; ms.gt_c1[g][3] = 0;
            ms.gt_commit[g][0] = ls.terms[q].cnt_off; ms.gt_commit[g][1] = 0; ms.gt_commit[g][2] = -1; ms.gt_commit[g][3] = 0;
            for (int j = 0; j < p.n_counters; j++)
              if (p.counters[j].topo_col >= 0 && p.counters[j].smem_off == ls.terms[q].cnt_off) {
                ms.gt_commit[g][1] = ls.cinfo[j].inc; ms.gt_commit[g][2] = ls.cinfo[j].pts_idx; ms.gt_commit[g][3] = ls.cinfo[j].n_present;
                ms.gt_c1[g][3] = p.counters[j].n_domains;
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
        ok &= has ? (c <= lt.lim + ms.relax[q]) : (lt.miss_rejects == 0);     // (relax > 0: closed cells close to reopening publish their nodes as dormant candidates)
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
    MPH_MARK(0);
    __syncthreads();                                                    // S1
    MPH_MARK(1);
    const unsigned long long tagbits = (unsigned long long)tag << KEY_TAG_SHIFT;
    const int par = (int)(wv & 1);
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
      // line layout (16 tagged words): [0] key 0  [1] key 7 | more-bit  [2..7] keys 1..6  [8..15] payloads 0..7 — the poller of a
      // line reads words 0-1 in one 16-byte load: the list's best key and its last key (with "this tile has more feasible nodes")
      unsigned long long kw = (unsigned long long)mykey;
      if (lane == MULTI_M - 1 && total > L) kw |= 1ull << MULTI_MORE_BIT;
      const int kpos = lane == 0 ? 0 : (lane == MULTI_M - 1 ? 1 : lane + 1);
      {   // (node shards too: the lines stay on this GPU; what crosses NVLink is one summary per rank, below)
        unsigned long long *myslots = p.slots + ((size_t)par * CCSIM_MAX_GRID + cta) * SLOT_STRIDE;
        if (lane < MULTI_M) {
          st_slot(&myslots[kpos], kw | tagbits);
          st_slot(&myslots[MULTI_M + lane], pay | tagbits);
        }
      }
    }
    MPH_MARK(2);
    // ---- gather, level 1: the lines of THIS GPU's CTAs. Every thread waits for its own (<= 2) entries — key word and payload word —
    //      so that the bar and the entries cost ONE L2 round trip after the slowest CTA's line lands ----
    const unsigned long long *lbase = p.slots + (size_t)par * CCSIM_MAX_GRID * SLOT_STRIDE;
    uint32_t tloc = 0u, kloc = 0u;
    unsigned long long ea[2] = {0ull, 0ull}, eb[2] = {0ull, 0ull};
    {
      const unsigned long long *pa[2], *pb[2];
      bool need[2];
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        const int e = tid + u * LEAN_THREADS, ee = e & (MULTI_M - 1);
        const unsigned long long *ln = lbase + (size_t)(e >> 3) * SLOT_STRIDE;
        pa[u] = ln + (ee == 0 ? 0 : (ee == MULTI_M - 1 ? 1 : ee + 1)); pb[u] = ln + MULTI_M + ee;
        need[u] = e < tot;
      }
      unsigned spins = 0;
      while (need[0] | need[1]) {
        #pragma unroll
        for (int u = 0; u < 2; u++) if (need[u]) { ea[u] = ld_slot(pa[u]); eb[u] = ld_slot(pb[u]); }
        #pragma unroll
        for (int u = 0; u < 2; u++) if (need[u] && (uint32_t)(ea[u] >> KEY_TAG_SHIFT) == tag && (uint32_t)(eb[u] >> KEY_TAG_SHIFT) == tag) need[u] = false;
        if (++spins > WATCHDOG_SPINS) { ms.dead = 1; if (need[0]) ea[0] = eb[0] = 0ull; if (need[1]) ea[1] = eb[1] = 0ull; break; }
      }
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        const int e = tid + u * LEAN_THREADS, ee = e & (MULTI_M - 1);
        if (e < tot) {
          if (ee == 0) kloc = max(kloc, (uint32_t)ea[u]);
          if (ee == MULTI_M - 1 && ((ea[u] >> MULTI_MORE_BIT) & 1ull)) tloc = max(tloc, (uint32_t)ea[u]);
        }
      }
    }
    tloc = __reduce_max_sync(0xffffffffu, tloc);
    kloc = __reduce_max_sync(0xffffffffu, kloc);
    if (lane == 0) { ms.red[warp] = tloc; ms.red2[warp] = kloc; }
    __syncthreads();                                                    // G1 (also: ms.dead)
    uint32_t Tlist = __reduce_max_sync(0xffffffffu, lane < LEAN_WARPS ? ms.red[lane] : 0u);
    uint32_t kbest = __reduce_max_sync(0xffffffffu, lane < LEAN_WARPS ? ms.red2[lane] : 0u);
    bool dead = ms.dead != 0;
    // The replay bar: T = the largest "last key" of a list whose tile has unseen feasible nodes is the lowest VALID bar; any
    // higher bar is valid too, just more conservative. The replay holds MULTI_CAP candidates, and it rarely needs more than the
    // best few dozen before a PTS minimum moves, so the bar is set `delta` below the best key (never below T); delta follows the
    // previous waves (doubled when the replay ran out of candidates above an artificial bar, shrunk when too many qualified).
    // Every CTA of every rank computes the same sequence from the same exchanged data.
    uint32_t T = max(Tlist, kbest > delta ? kbest - delta : 0u);
    // ---- the entries keyed >= T go into shared memory (unordered; keys are unique) ----
    int C = 0;
    for (int pass = 0; pass < 2 && !dead; pass++) {
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t ck = (uint32_t)ea[u];
        if (u * LEAN_THREADS < tot) multi_append(ck != 0u && ck >= T, ck, eb[u], lane);      // (warp-uniform guard: warps beyond the entries skip)
      }
      __syncthreads();                                                  // G2
      C = ms.ncand;
      if (C <= MULTI_CAP || pass == 1) break;
      // More candidates than the replay holds: raise T to the lowest of 64 equal steps between T and the best key that leaves
      // <= MULTI_CAP candidates; every CTA sees the same data and decides alike.
      if (tid < MULTI_BINS) ms.hist[tid] = 0u;
      __syncthreads();
      const unsigned long long range = (unsigned long long)(kbest - T) + 1ull;
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint32_t ck = (uint32_t)ea[u];
        if (ck != 0u && ck >= T) atomicAdd(&ms.hist[(unsigned)(((unsigned long long)(ck - T) * MULTI_BINS) / range)], 1u);
      }
      __syncthreads();
      int bsel = MULTI_BINS;
      { unsigned sum = 0; for (int bq = MULTI_BINS - 1; bq >= 0; bq--) { sum += ms.hist[bq]; if (sum > (unsigned)MULTI_CAP) break; bsel = bq; } }
      T = (bsel >= MULTI_BINS) ? kbest : T + (uint32_t)(((unsigned long long)bsel * range + (MULTI_BINS - 1)) / MULTI_BINS);
      __syncthreads();
      if (tid == 0) ms.ncand = 0;
      if (cta == 0 && tid == 0) ms.st_overflow++;
      __syncthreads();
    }
    if (XGPU) {
      // ---- gather, level 2 (node shards): every rank now holds ITS candidates keyed >= its bar T_r (<= MULTI_CAP of them, the same in
      //      all of its CTAs). One summary per (source, destination) pair crosses NVLink — {best key, count, T_r, T_list_r, the
      //      candidates} written by ONE CTA of the source as a few 128-byte stores — instead of every CTA's line going to every
      //      rank and world x grid lines being polled by every CTA. The global bar max(max_r T_r, best - delta) is at least every
      //      rank's own bar, so the union of the summaries holds every node keyed above it: same candidates as one GPU would see. ----
      const int xpar = (int)((wv + p.xwave0) & 1);
      const int Cl = C < MULTI_CAP ? C : MULTI_CAP;
      uint32_t lkey = 0u; unsigned long long lpay = 0ull;        // this rank's candidate `tid`, kept across the shared-memory rebuild
      if (tid < Cl) { lkey = ms.ckey[tid]; lpay = (unsigned long long)ms.cdom[tid] | ((unsigned long long)ms.cnext[tid] << MULTI_NEXT_SHIFT); }
      if (!dead)
        for (int d = cta; d < p.world; d += p.grid) {            // CTA d of the source writes the copy for rank d
          if (d == p.rank) continue;
          unsigned long long *dst = p.xslots_peer[d] + XLINES_OFF + ((size_t)xpar * CCSIM_MAX_WORLD + p.rank) * CCSIM_MAX_GRID * SLOT_STRIDE;
          if (tid < 4 + 2 * Cl) {
            unsigned long long v;
            if (tid == 0) v = (unsigned long long)kbest | ((unsigned long long)Cl << 32);
            else if (tid == 1) v = T;
            else if (tid == 2) v = Tlist;
            else if (tid == 3) v = 0ull;
            else { const int ci = (tid - 4) >> 1; v = (tid & 1) ? ((unsigned long long)ms.cdom[ci] | ((unsigned long long)ms.cnext[ci] << MULTI_NEXT_SHIFT)) : (unsigned long long)ms.ckey[ci]; }
            st_slot_sys(dst + tid, v | tagbits);
          }
        }
      __syncthreads();                                                  // X0: the local candidates are in registers / on their way
      if (tid == 0) ms.ncand = 0;
      uint32_t xkb = kbest, xtl = Tlist, xbar = T;
      if (tid < p.world) {
        int cr = 0;
        if (tid != p.rank && !dead) {
          const unsigned long long *src = p.xslots_peer[p.rank] + XLINES_OFF + ((size_t)xpar * CCSIM_MAX_WORLD + tid) * CCSIM_MAX_GRID * SLOT_STRIDE;
          unsigned long long a, b, c2;
          unsigned spins = 0;
          for (;;) {
            ld_line2<true>(src, a, b); c2 = ld_slot_sys(src + 2);
            if ((uint32_t)(a >> KEY_TAG_SHIFT) == tag && (uint32_t)(b >> KEY_TAG_SHIFT) == tag && (uint32_t)(c2 >> KEY_TAG_SHIFT) == tag) break;
            if (++spins > WATCHDOG_SPINS) { ms.dead = 1; a = b = c2 = 0ull; break; }
          }
          xkb = (uint32_t)a; cr = (int)((a >> 32) & 0xffu); xbar = (uint32_t)b; xtl = (uint32_t)c2;
          if (cr > MULTI_CAP) cr = MULTI_CAP;
        }
        ms.xcount[tid] = cr;
      }
      if (warp == 0) {       // (world <= 32: the pollers are lanes of warp 0; the other lanes carry this rank's own values)
        xkb = __reduce_max_sync(0xffffffffu, xkb); xtl = __reduce_max_sync(0xffffffffu, xtl); xbar = __reduce_max_sync(0xffffffffu, xbar);
        if (lane == 0) { ms.xglob[0] = xkb; ms.xglob[1] = xtl; ms.xglob[2] = xbar; }
      }
      __syncthreads();                                                  // X1
      dead = ms.dead != 0;
      kbest = ms.xglob[0]; Tlist = ms.xglob[1];
      T = max(ms.xglob[2], kbest > delta ? kbest - delta : 0u);
      // the other ranks' candidates: entry e of the concatenated summaries -> (rank, index); <= 2 per thread
      uint32_t rkey[MULTI_XPT]; unsigned long long rpay[MULTI_XPT];
      #pragma unroll
      for (int u = 0; u < MULTI_XPT; u++) {
        rkey[u] = 0u; rpay[u] = 0ull;
        int e = tid + u * LEAN_THREADS, r = 0;
        while (r < p.world && e >= ms.xcount[r]) { e -= ms.xcount[r]; r++; }
        if (r < p.world && !dead) {
          const unsigned long long *src = p.xslots_peer[p.rank] + XLINES_OFF + ((size_t)xpar * CCSIM_MAX_WORLD + r) * CCSIM_MAX_GRID * SLOT_STRIDE + 4 + 2 * e;
          unsigned long long a, b;
          unsigned spins = 0;
          for (;;) {
            ld_line2<true>(src, a, b);
            if ((uint32_t)(a >> KEY_TAG_SHIFT) == tag && (uint32_t)(b >> KEY_TAG_SHIFT) == tag) break;
            if (++spins > WATCHDOG_SPINS) { ms.dead = 1; a = b = 0ull; break; }
          }
          rkey[u] = (uint32_t)a; rpay[u] = b;
        }
      }
      for (int pass = 0; pass < 2; pass++) {
        multi_append(lkey != 0u && lkey >= T, lkey, lpay, lane);
        #pragma unroll
        for (int u = 0; u < MULTI_XPT; u++) multi_append(rkey[u] != 0u && rkey[u] >= T, rkey[u], rpay[u], lane);
        __syncthreads();                                                // X2
        C = ms.ncand;
        if (C <= MULTI_CAP || pass == 1) break;
        if (tid < MULTI_BINS) ms.hist[tid] = 0u;
        __syncthreads();
        const unsigned long long range = (unsigned long long)(kbest - T) + 1ull;
        if (lkey != 0u && lkey >= T) atomicAdd(&ms.hist[(unsigned)(((unsigned long long)(lkey - T) * MULTI_BINS) / range)], 1u);
        #pragma unroll
        for (int u = 0; u < MULTI_XPT; u++)
          if (rkey[u] != 0u && rkey[u] >= T) atomicAdd(&ms.hist[(unsigned)(((unsigned long long)(rkey[u] - T) * MULTI_BINS) / range)], 1u);
        __syncthreads();
        int bsel = MULTI_BINS;
        { unsigned sum = 0; for (int bq = MULTI_BINS - 1; bq >= 0; bq--) { sum += ms.hist[bq]; if (sum > (unsigned)MULTI_CAP) break; bsel = bq; } }
        T = (bsel >= MULTI_BINS) ? kbest : T + (uint32_t)(((unsigned long long)bsel * range + (MULTI_BINS - 1)) / MULTI_BINS);
        __syncthreads();
        if (tid == 0) ms.ncand = 0;
        if (cta == 0 && tid == 0) ms.st_overflow++;
        __syncthreads();
      }
      dead = dead || ms.dead != 0;
    }
    const bool overflowed = C > MULTI_CAP || T > max(Tlist, kbest > delta ? kbest - delta : 0u);
    if (C > MULTI_CAP) C = MULTI_CAP;     // (cannot happen after the second pass: the raised T admits <= MULTI_CAP keys)
    if (cta == 0 && tid == 0) ms.st_cand += C;
    MPH_MARK(3);
    // ---- replay: the reference cycles k, k+1, ... this wave can decide; every CTA does the same, in ONE warp and without a
    //      barrier: MULTI_CPT candidates per lane in registers, lane q < n_gt also owns counter term q (its constants, and the
    //      minimum / multiplicity of the PTS constraint it tracks, stay in registers for the whole wave) ----
    if (warp == 0) {
      int32_t acc = 0;
      bool ran_dry = false, any_relax = false;
      if (!dead && !ms.dead) {
        const int n_gt = ms.n_gt;
        uint32_t ck[MULTI_CPT], cd[MULTI_CPT];
        uint32_t second = 0u;
        #pragma unroll
        for (int j = 0; j < MULTI_CPT; j++) {
          const int idx = j * 32 + lane;
          ck[j] = 0u; cd[j] = 0u;
          if (idx < C) { ck[j] = ms.ckey[idx]; cd[j] = ms.cdom[idx]; }
        }
        int4 gc = make_int4(0, 0, -1, 0), c1 = make_int4(0, 0, 0, 0);
        int32_t my_min = 0, my_num = 0;
        if (lane < n_gt) {
          gc = *reinterpret_cast<const int4 *>(&ms.gt_commit[lane][0]);   // {cnt_off, inc, pts_idx, n_present}
          c1 = *reinterpret_cast<const int4 *>(&ms.gt_c1[lane][0]);       // {lim, shift, mask, n_domains}
          if (gc.z >= 0) { my_min = ls.ptsmin[gc.z]; my_num = ls.ptsnum[gc.z]; }
        }
        uint32_t fmask[MULTI_GT];                    // payload field mask of term q, pre-shifted (0: no such term)
        #pragma unroll
        for (int q = 0; q < MULTI_GT; q++) fmask[q] = q < n_gt ? ((uint32_t)ms.gt_c1[q][2] << ms.gt_c1[q][1]) : 0u;
        const int32_t lim_off = c1.x - my_min;       // PTS: maxSkew - selfMatch (the limit follows the global minimum)
        bool lim_moved = false;
        const bool single_use = ms.single_use != 0;
        // commits this wave may still decide: --max-limit (simulator.go:300-305), the output capacity, MULTI_MAX_ACC
        long long room = p.pod_cap - k;
        if (p.max_pods > 0 && p.max_pods - k < room) room = p.max_pods - k;
        const int32_t acc_limit = room < MULTI_MAX_ACC ? (int32_t)room : MULTI_MAX_ACC;
        int rounds = 0;
        bool ended_by_rescan = false;
        // ---- dormant candidates. A PTS term whose limit is about to move (few domains left at the global minimum) was scanned with
        //      a look-ahead (ms.relax): nodes in cells up to MULTI_RELAX_R over the limit are in the lists too. They cannot win while
        //      their cell is over the limit (dormant: key 0 in ck[], like a dead candidate) and wake up when a minimum move lifts the
        //      limit over their cell — the wave then goes on instead of ending for a rescan. It still has to end when the limit
        //      reaches a cell that was NOT published: `unpub_min` = the smallest count above limit + look-ahead at the start of the
        //      wave (such cells are closed, so their counts stand for the whole wave).
        //      The round loop itself is unchanged: a candidate whose cell fills is zeroed as before. Waking is a REBUILD (rare: at the
        //      start of a look-ahead wave and after a minimum move of a look-ahead term): every slot's key is read again from the
        //      wave's candidate arrays — first-life key, or second-life key when bit j of `second` says the slot has won already
        //      (single-use templates: gone) — and zeroed when one of its cells is over the limit as counters and limits stand now. ----
        int32_t rlx = 0, unpub_min = INT32_MAX;
        if (lane < n_gt) rlx = lds_s32(MS_SA(relax) + 4u * (uint32_t)lds_s32(MS_SA(gt_term) + 4u * (uint32_t)lane));
        any_relax = __any_sync(0xffffffffu, rlx > 0);
        bool need_rebuild = any_relax;
        if (any_relax) {
          for (unsigned nm = __ballot_sync(0xffffffffu, rlx > 0); nm; nm &= nm - 1) {
            const int q = __ffs(nm) - 1;
            const int32_t off = __shfl_sync(0xffffffffu, gc.x, q), ndom = __shfl_sync(0xffffffffu, c1.w, q);
            const int32_t publim = __shfl_sync(0xffffffffu, c1.x + rlx, q);
            const uint32_t ca = cnt_sa + 4u * (uint32_t)off;
            int32_t m = INT32_MAX;
            #pragma unroll 1
            for (int d = lane; d < ndom; d += 32) { const int32_t c = lds_s32(ca + 4u * d); if (c > publim) m = min(m, c); }
            m = __reduce_min_sync(0xffffffffu, m);
            if (lane == q) unpub_min = m;
          }
          if (cta == 0 && lane == 0) ms.st_relaxed++;
        }
        if (cta == 0 && lane == 0) ms.ph[6] += clock64() - ms.tc0;      // replay set-up
        for (;;) {
          if (need_rebuild) {       // (one call site, off the round's critical path; see "dormant candidates")
            need_rebuild = false;
            __syncwarp();           // the counter cells / limits written by the term lanes, before every lane reads them
            uint32_t badm = 0u;
            #pragma unroll 1
            for (int q = 0; q < n_gt; q++) {
              const int4 tg = *reinterpret_cast<const int4 *>(&ms.gt_commit[q][0]);   // {cnt_off, inc, pts_idx, n_present}
              const int4 tc = *reinterpret_cast<const int4 *>(&ms.gt_c1[q][0]);       // {lim (kept current by lane q), shift, mask, n_domains}
              #pragma unroll
              for (int j = 0; j < MULTI_CPT; j++) {
                const int32_t v = (int32_t)((cd[j] >> tc.y) & (uint32_t)tc.z) - 1;
                if (v >= 0 && lds_s32(cnt_sa + 4u * (uint32_t)(tg.x + v)) > tc.x) badm |= 1u << j;
              }
            }
            #pragma unroll
            for (int j = 0; j < MULTI_CPT; j++) {
              const int idx = j * 32 + lane;
              uint32_t base = 0u;
              if (idx < C && !((badm >> j) & 1u)) {
                const uint32_t k0 = (uint32_t)lds_s32(MS_SA(ckey) + 4u * (uint32_t)idx);
                if (!((second >> j) & 1u)) base = k0;
                else if (!single_use) { const uint32_t ns = (uint32_t)lds_s32(MS_SA(cnext) + 4u * (uint32_t)idx); base = ns ? ((ns << MULTI_IDX_BITS) | (k0 & MULTI_IDX_MASK)) : 0u; }
              }
              ck[j] = base;
            }
          }
          rounds++;
          uint32_t m = ck[0];
          #pragma unroll
          for (int j = 1; j < MULTI_CPT; j++) m = max(m, ck[j]);
          const uint32_t g = __reduce_max_sync(0xffffffffu, m);
          if (g == 0u || g < T) { ran_dry = true; break; }      // nothing left, or an unseen node could rank above g
          // ---- commit pod k+acc (assume -> AssumePod -> NodeInfo.update(+1): schedule_one.go:967-984, types.go:409-427) ----
          // the owner lane contributes the winner's payload (keys are unique: exactly one lane and slot match)
          uint32_t dsel = 0u;
          #pragma unroll
          for (int j = 0; j < MULTI_CPT; j++) dsel = (ck[j] == g) ? cd[j] : dsel;
          const uint32_t pay = __reduce_or_sync(0xffffffffu, dsel);
          // the winner comes back once with the key it has after this clone (if it still fits); when a node wins for the second
          // time in a wave its third key is unknown: the wave ends after that commit
          bool sec = false;
          if (single_use) {           // a clone blocks its own node (hostname anti-affinity): the winner just leaves
            #pragma unroll
            for (int j = 0; j < MULTI_CPT; j++) { const bool w = ck[j] == g; ck[j] = w ? 0u : ck[j]; second |= (uint32_t)w << j; }   // (bit j: this slot has won — a rebuild leaves it out)
          } else {
            #pragma unroll
            for (int j = 0; j < MULTI_CPT; j++)
              if (ck[j] == g) {
                sec = (second >> j) & 1u;
                const uint32_t ns = sec ? 0u : (uint32_t)lds_s32(MS_SA(cnext) + 4u * (uint32_t)(j * 32 + lane));
                ck[j] = ns ? ((ns << MULTI_IDX_BITS) | (g & MULTI_IDX_MASK)) : 0u;
                second |= 1u << j;
              }
          }
          // Only what the next round depends on happens here: the counter cells of the winner's domains (lane q = term q; the
          // host guarantees one term per incremented replicated counter), whether a cell went over its limit, whether a PTS
          // minimum moved. The winner's row (NodeInfo.update, node-local counters) is brought up to date by its own thread after the wave.
          bool minchg = false;
          uint32_t fullf = 0u;
          if (lane < n_gt) {
            const int32_t v = (int32_t)((pay >> c1.y) & (uint32_t)c1.z) - 1;
            if (v >= 0) {
              const uint32_t ca = cnt_sa + 4u * (uint32_t)(gc.x + v);
              const int32_t old = lds_s32(ca), nv = old + gc.y;
              sts_s32(ca, nv);
              if (nv > c1.x) fullf = (uint32_t)(v + 1) << c1.y;            // candidates in this cell are dead from now on
              if (gc.y && gc.z >= 0 && v < gc.w && old == my_min) { my_num--; minchg = my_num <= 0; }   // the global minimum of this constraint moves: limits change, rescan
            }
          }
          if (lane == 0) sts_s32(MS_SA(acc_node) + 4u * (uint32_t)acc, ckey_index(g));
          acc++;
          // A PTS minimum moved (filtering.go:56-69: minMatchNum): recount it and move the term's limit. The wave goes on unless
          // the move changes some node's feasibility: that takes a domain whose count lies in (old limit, new limit] — nodes there
          // were rejected by the scan (or killed earlier in this wave) and pass now. Without such a domain every verdict so far
          // stands (the 8-region constraint of C4 moves its minimum every 8 placements and never binds).
          bool rescan = false, woke = false;
          const unsigned mc = __ballot_sync(0xffffffffu, minchg);
          for (unsigned nm = mc; nm; nm &= nm - 1) {
            const int q = __ffs(nm) - 1;
            const int32_t off = __shfl_sync(0xffffffffu, gc.x, q), npres = __shfl_sync(0xffffffffu, gc.w, q), ndom = __shfl_sync(0xffffffffu, c1.w, q);
            const int32_t lim_old = __shfl_sync(0xffffffffu, c1.x, q), loff = __shfl_sync(0xffffffffu, lim_off, q);
            const uint32_t ca = cnt_sa + 4u * (uint32_t)off;
            int32_t mn = INT32_MAX;
            #pragma unroll 1
            for (int d = lane; d < npres; d += 32) mn = min(mn, lds_s32(ca + 4u * d));
            mn = __reduce_min_sync(0xffffffffu, mn);
            const long long liml = (long long)loff + (long long)mn;
            const int32_t lim_new = liml > INT32_MAX ? INT32_MAX : (liml < INT32_MIN ? INT32_MIN : (int32_t)liml);
            int32_t num = 0;
            bool hit = false;
            #pragma unroll 1
            for (int d = lane; d < ndom; d += 32) { const int32_t c = lds_s32(ca + 4u * d); num += (d < npres) & (c == mn); hit |= (c > lim_old) & (c <= lim_new); }
            num = __reduce_add_sync(0xffffffffu, num);
            // a term scanned with look-ahead published the nodes of its closed cells: the move only matters when the new limit
            // reaches a cell that was not published; the cells in (old limit, new limit] wake their candidates up instead
            const bool rq = __shfl_sync(0xffffffffu, rlx, q) > 0;
            if (rq) { rescan |= (lim_new >= __shfl_sync(0xffffffffu, unpub_min, q)) | (p.debug_flags & 1); woke = true; }
            else rescan |= __any_sync(0xffffffffu, hit) | (p.debug_flags & 1);
            if (lane == q) { my_min = mn; my_num = num; c1.x = lim_new; lim_moved = true; sts_s32(MS_SA(gt_c1) + 16u * (uint32_t)q, lim_new); }
          }
          need_rebuild = woke & !rescan;
          // (no statistics, special registers or kernel parameters are touched inside the round loop: one S2R on this dependent
          //  chain costs as much as ten ALU instructions)
          bool stopb = __any_sync(0xffffffffu, sec) | rescan;
          if (stopb | (acc >= acc_limit)) { ended_by_rescan = rescan; break; }
          // only the candidates sitting in a counter cell that this commit pushed over its limit die (monotone: for the rest of
          // the wave); the fields of different terms are disjoint bit ranges, so one OR-reduction carries all filled cells
          // (the rebuild at the top of the next round sets every candidate as counters and limits stand by then — `fullf` was taken
          //  against the limits before the move)
          const uint32_t F = woke ? 0u : __reduce_or_sync(0xffffffffu, fullf);
          if (F) {
            #pragma unroll
            for (int q = 0; q < MULTI_GT; q++) {
              const uint32_t f = F & fmask[q];
              if (f) {
                #pragma unroll
                for (int j = 0; j < MULTI_CPT; j++)
                  if ((cd[j] & fmask[q]) == f) ck[j] = 0u;     // (dead — in a look-ahead wave: dormant, a rebuild may bring it back)
              }
            }
          }
        }
        if (cta == 0 && lane == 0) { ms.st_rounds += rounds; if (ended_by_rescan) ms.ph[7] += 1; }      // (ph[7]: waves ended by a minimum move that changes verdicts)
        // ---- the next wave's look-ahead, per PTS term on a replicated counter: its limit is about to move (<= MULTI_RELAX_K present
        //      domains left at the minimum) and the closed domains are not the majority (their nodes would crowd the live ones out of
        //      the tiles' top-M lists). A look-ahead wave that could not place anything is repeated strictly. Every CTA of every rank
        //      decides alike (same counters, same replay). ----
        {
          const bool strict_next = (acc == 0 && any_relax) || (p.debug_flags & 16u);
          const bool elig = lane < n_gt && gc.z >= 0 && gc.y > 0 && my_num <= MULTI_RELAX_K && c1.x < INT32_MAX - 2 * MULTI_RELAX_R && !strict_next;
          int32_t nrl = 0;
          for (unsigned nm = __ballot_sync(0xffffffffu, elig); nm; nm &= nm - 1) {
            const int q = __ffs(nm) - 1;
            const int32_t off = __shfl_sync(0xffffffffu, gc.x, q), npres = __shfl_sync(0xffffffffu, gc.w, q), lim = __shfl_sync(0xffffffffu, c1.x, q);
            const uint32_t ca = cnt_sa + 4u * (uint32_t)off;
            int32_t closed = 0;
            #pragma unroll 1
            for (int d = lane; d < npres; d += 32) closed += lds_s32(ca + 4u * d) > lim;
            closed = __reduce_add_sync(0xffffffffu, closed);
            if (lane == q && 2 * closed <= npres) nrl = MULTI_RELAX_R;
          }
          if (p.debug_flags & 32u) nrl = (lane < n_gt && gc.z >= 0 && gc.y > 0 && c1.x < INT32_MAX - 2 * MULTI_RELAX_R && !strict_next) ? MULTI_RELAX_R : 0;   // tests: look-ahead on every PTS term, every wave
          if (lane < n_gt) sts_s32(MS_SA(relax) + 4u * (uint32_t)lds_s32(MS_SA(gt_term) + 4u * (uint32_t)lane), nrl);
          if (cta == 0 && lane == 0 && acc == 0 && any_relax) ms.st_empty++;
        }
        // the limits that moved go back to the Filter constants of the next scan
        if (lim_moved) { ls.terms[ms.gt_term[lane]].lim = c1.x; ms.gt_c1[lane][0] = c1.x; }
        if (lane < n_gt && gc.z >= 0) { ls.ptsmin[gc.z] = my_min; ls.ptsnum[gc.z] = my_num; }
      }
      if (lane == 0 && cta == 0 && (p.debug_flags & 4))
        printf("wave %lld k=%lld acc=%d C=%d T=%08x Tlist=%08x kbest=%08x delta=%08x ran_dry=%d look_ahead=%d first=%d last=%d\n", wv, k, acc, C, T, Tlist, kbest, delta,
               (int)ran_dry, (int)any_relax, acc ? ms.acc_node[0] : -1, acc ? ms.acc_node[acc - 1] : -1);
      if (lane == 0) {
        ms.accepted = acc; ms.ncand = 0;
        // next wave's bar distance: the replay ran out of candidates above a bar that was higher than it had to be -> look further
        // down next time; far more candidates than a wave uses -> look less far
        uint32_t nd = delta;
        if (ran_dry && T > Tlist) nd = delta < (1u << 30) ? delta * 2u : delta;
        else if (overflowed) nd = max(delta / 2u, 1u << 8);
        else if (!ran_dry && C > MULTI_CAP / 2) nd = max(delta - delta / 8u, 1u << 8);
        ms.delta = nd;
        if (dead || ms.dead) ls.stop = 3;
        else if (acc == 0 && !any_relax) ls.stop = 1;          // no feasible node anywhere: the pod is unschedulable (after a look-ahead wave: rescan strictly first)
      }
    }
    __syncthreads();                                                    // R: the accepted list, counters, limits
    MPH_MARK(4);
    const int32_t acc = ms.accepted;
    // ClusterCapacityBinder.Bind + postBindHook: pod k+i -> node (plugin.go:34-53; simulator.go:297-312). Every CTA knows the
    // whole list; CTA 0 (of every rank: each keeps the whole sequence) records it.
    if (cta == 0 && tid < acc && k + tid < p.pod_cap) p.pod_node[k + tid] = ms.acc_node[tid];
    // ---- assume -> AssumePod -> NodeInfo.update(+1) (schedule_one.go:967-984, types.go:409-427): every thread looks for its own
    //      node among the winners (a node is accepted at most twice per wave). Only this thread reads the row before the next S1. ----
    if (tid < cnt_nodes && acc) {
      const int32_t mynode = p.node_base + lo + tid;
      int mult = 0;
      for (int i = 0; i < acc; i++) mult += (ms.acc_node[i] == mynode);
      if (mult) {
        const ccsim_template &t = ls.tmpl;
        const int32_t jw = tid;
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
    MPH_MARK(5);
    k += acc;
    delta = ms.delta;
    if (ls.stop) break;
    for (int c = 0; c < ls.tmpl.n_pts; c++)
      if (!ls.tmpl.pts[c].min_zero && (ls.ptsnum[c] <= 0 || (p.debug_flags & 2)) && p.counters[ls.tmpl.pts[c].counter].n_present > 0) lean_pts_recount(p, smem_cnt, c);
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
      for (int q = 0; q < 8; q++) o->phase_cycles[q] = ms.ph[q];
      o->stat[0] = ms.st_cand; o->stat[1] = ms.st_overflow; o->stat[2] = ms.st_rounds;
      if (p.debug_flags & 8u)
        printf("multi-commit replay: waves %lld rounds %lld look-ahead waves %d (without a placement: %d) | cycles: replay %lld set-up %lld\n",
               limit_hit ? wv : wv + 1, ms.st_rounds, ms.st_relaxed, ms.st_empty, ms.ph[4], ms.ph[6]);
    }
  }
}
