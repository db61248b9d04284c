// ccsim_device.cuh — device-side data layout and the per-node predicate / score functions (sm_100a).
//
// One predicate-eval = one (pod attempt, node) pair through eval_node(): what checkNode does once in the reference
// (vendor/k8s.io/kubernetes/pkg/scheduler/schedule_one.go:644-666 -> framework/runtime/framework.go:897-930), followed
// for feasible nodes by the node-local part of RunScorePlugins (framework.go:1137-1244).
//
// Integer/indexing work only: int64 compares, one exact int64 quotient in [0,100] per scored resource
// (least_allocated.go:52-61), IEEE float64 for BalancedAllocation (balanced_allocation.go:146-180) with explicit
// round-to-nearest intrinsics so that nothing is contracted into an FMA.
#pragma once
#include <stdint.h>
#include "../../include/ccsim.h"

#define CCSIM_MAX_GRID 160  /* >= SM count of the part (B200: 148) */

struct DevCounter {
  int32_t topo_col;   // -1: node-local column
  int32_t n_domains;
  int32_t n_present;
  int32_t inc;
  int32_t smem_off;   // offset (in int32) into the CTA's shared counter area, or -1: per-CTA replica in global memory
  int32_t is_aff;     // counter belongs to a required pod-affinity key
  int32_t *init;      // [n_domains] device copy of the initial counts (node-local: snapshot column)
  int32_t *work;      // node-local: working column [n]; replicated-global: base of grid*n_domains replicas
};

struct DevOut {
  int64_t placed;
  int32_t stop_code;
  int32_t error;      // 0 ok, 1 watchdog (a CTA never saw its peers' slots), 2 pod_node overflow
  int64_t waves;
  int64_t evals;
  int32_t ptsmin[CCSIM_MAX_PTS];   // global minima at the terminal cycle (for the diagnosis pass)
  int64_t aff_total;
  unsigned long long reason_hist[CCSIM_R_TOTAL];
  unsigned long long preempt_no_victims;
  unsigned long long n_diag;
};

struct DevParams {
  int32_t n;            // nodes of this shard
  int32_t n_global;     // nodes of the whole cluster
  int32_t node_base;    // global index of local node 0
  int32_t n_scalars, taint_words, static_words, n_topo, n_templates, n_counters, n_classes;
  int32_t grid;         // CTAs of the persistent kernel
  int32_t chunk;        // nodes per CTA (contiguous ownership)
  int32_t rank, world;
  // immutable columns
  const int64_t *alloc_cpu, *alloc_mem, *alloc_eph;
  const int32_t *alloc_pods;
  const int64_t *alloc_scalar[CCSIM_MAX_SCALARS];
  const uint64_t *taint_mask, *static_mask;
  const int32_t *topo[CCSIM_MAX_TOPO_COLS];
  // mutable working columns (restored from the snapshot copies before each run)
  int64_t *req_cpu, *req_mem, *req_eph, *nz_cpu, *nz_mem;
  int32_t *npods;
  int64_t *req_scalar[CCSIM_MAX_SCALARS];
  uint64_t *placed_mask;   // nullptr unless a template has hostPorts
  uint64_t taint_nosched[CCSIM_MAX_TAINT_WORDS], taint_prefer[CCSIM_MAX_TAINT_WORDS];
  const ccsim_template *templates;
  DevCounter counters[CCSIM_MAX_COUNTERS];
  int32_t *final_cnt;      // concatenated final replicated counters (written by CTA 0 at exit)
  int32_t final_off[CCSIM_MAX_COUNTERS];
  // exchange: slots[parity][cta][class]
  unsigned long long *slots;
  // cross-GPU exchange (multi-GPU persistent mode): xslots[parity][rank][class] in every peer's memory
  unsigned long long *xslots_peer[8];
  int32_t *pod_node;
  int64_t pod_cap;
  int64_t max_pods;
  DevOut *out;
  const int32_t *taint_list_off;
  const uint8_t *taint_list;
};

// ---- key packing ----------------------------------------------------------------------------------------------
// [63:52] tag (wave+1, 12 bit, never 0)  [51:32] score+1 (0 = no feasible node)  [31:0] 0xFFFFFFFF - global node index
// max over keys = highest score, ties -> lowest node index = "first max in scan order" (selectHost, schedule_one.go:894-941).
#define KEY_TAG_SHIFT 52
#define KEY_BODY_MASK ((1ull << KEY_TAG_SHIFT) - 1)
__device__ __forceinline__ unsigned long long pack_key(int64_t score, uint32_t gidx) {
  return ((unsigned long long)(score + 1) << 32) | (unsigned long long)(0xFFFFFFFFu - gidx);
}
__device__ __forceinline__ int64_t key_score(unsigned long long k) { return (int64_t)((k & KEY_BODY_MASK) >> 32) - 1; }
__device__ __forceinline__ uint32_t key_index(unsigned long long k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu); }

// ---- exact scorers --------------------------------------------------------------------------------------------
// leastRequestedScore: ((capacity - requested) * 100) / capacity, truncating int64 (least_allocated.go:52-61).
// The quotient is in [0,100]: estimate in fp32, then repair with one exact int64 remainder test (no 64-bit divide).
__device__ __forceinline__ int64_t least_requested_score(int64_t requested, int64_t capacity) {
  if (capacity == 0) return 0;
  if (requested > capacity) return 0;
  const int64_t x100 = (capacity - requested) * 100;
  int64_t q = (int64_t)(__fdividef((float)x100, (float)capacity));
  int64_t r = x100 - q * capacity;
  if (r < 0) { q -= 1; r += capacity; if (r < 0) { q = x100 / capacity; } }
  else if (r >= capacity) { q += 1; r -= capacity; if (r >= capacity) { q = x100 / capacity; } }
  return q;
}

// Fit.Score with LeastAllocated over cpu,mem: resource_allocation.go:48-114 + least_allocated.go:30-48
__device__ __forceinline__ int64_t score_least(int64_t a_cpu, int64_t a_mem, int64_t nz_cpu, int64_t nz_mem,
                                               const ccsim_template &t) {
  int64_t node_score = 0, wsum = 0;
  if (a_cpu != 0) { node_score += least_requested_score(nz_cpu + t.least_cpu, a_cpu) * t.least_w_cpu; wsum += t.least_w_cpu; }
  if (a_mem != 0) { node_score += least_requested_score(nz_mem + t.least_mem, a_mem) * t.least_w_mem; wsum += t.least_w_mem; }
  if (wsum == 0) return 0;
  if (wsum == 2) return node_score >> 1;   // both weights 1 (default): node_score >= 0
  return node_score / wsum;
}

// balancedResourceScorer over cpu,mem (balanced_allocation.go:146-180): float64, one rounding per operation.
__device__ __forceinline__ int64_t score_balanced(int64_t a_cpu, int64_t a_mem, int64_t r_cpu, int64_t r_mem,
                                                  const ccsim_template &t) {
  double f0 = 0.0, f1 = 0.0;
  int nf = 0;
  if (a_cpu != 0) {
    double fr = __ddiv_rn((double)(r_cpu + t.bal_cpu), (double)a_cpu);
    if (fr > 1.0) fr = 1.0;
    f0 = fr; nf = 1;
  }
  if (a_mem != 0) {
    double fr = __ddiv_rn((double)(r_mem + t.bal_mem), (double)a_mem);
    if (fr > 1.0) fr = 1.0;
    if (nf == 0) f0 = fr; else f1 = fr;
    nf++;
  }
  double sd = 0.0;
  if (nf == 2) sd = fabs(__dmul_rn(__dsub_rn(f0, f1), 0.5));   // (f0-f1)/2: exact scaling by a power of two
  return (int64_t)__dmul_rn(__dsub_rn(1.0, sd), 100.0);
}

// ---- per-CTA view of the dynamic cross-node state -------------------------------------------------------------
struct CtaState {
  int32_t *smem_cnt;                // replicated counters (shared memory area)
  int32_t ptsmin[CCSIM_MAX_PTS];    // (in shared memory) global minimum per PTS constraint
  int32_t ptsnum[CCSIM_MAX_PTS];    // number of present domains at the minimum
  long long aff_total;
};

__device__ __forceinline__ const int32_t *counter_base(const DevParams &p, int j, const int32_t *smem_cnt) {
  const DevCounter &c = p.counters[j];
  if (c.topo_col < 0) return c.work;
  if (c.smem_off >= 0) return smem_cnt + c.smem_off;
  return c.work + (size_t)blockIdx.x * c.n_domains;
}

// status codes for the diagnosis pass
#define ST_OK 0
#define ST_UNSCHEDULABLE 1
#define ST_UNRESOLVABLE 2

// Hot path: is node i (local index) feasible for template t, and if so its class (raw PreferNoSchedule intolerable
// count) and node-local score. Returns false if any enabled Filter plugin rejects the node.
__device__ __forceinline__ bool eval_node(const DevParams &p, const ccsim_template &t, const int32_t *smem_cnt,
                                          const int32_t *ptsmin, long long aff_total, int32_t i,
                                          int &cls, int64_t &score) {
  const int32_t n = p.n;
  // -- loads issued up front (coalesced: consecutive threads -> consecutive nodes) --
  const uint64_t taint0 = p.taint_mask[i];
  const int32_t a_pods = p.alloc_pods[i];
  const int32_t npods = p.npods[i];
  const int64_t a_cpu = p.alloc_cpu[i], a_mem = p.alloc_mem[i];
  const int64_t r_cpu = p.req_cpu[i], r_mem = p.req_mem[i];
  const int64_t z_cpu = p.nz_cpu[i], z_mem = p.nz_mem[i];
  bool ok = true;

  if (t.flags & CCSIM_TF_PREFILTER_NODES) {
    const int b = t.prefilter_bit;
    ok &= (bool)((p.static_mask[(size_t)(b >> 6) * n + i] >> (b & 63)) & 1ull);
  }
  // NodeUnschedulable (node_unschedulable.go:133-150)
  if (t.filter_enable & CCSIM_PL_NODE_UNSCHEDULABLE)
    ok &= !(((taint0 >> CCSIM_TAINT_UNSCHEDULABLE_BIT) & 1ull) && !(t.flags & CCSIM_TF_TOLERATES_UNSCHEDULABLE));
  // NodeName (node_name.go:81-83)
  if ((t.filter_enable & CCSIM_PL_NODE_NAME) && t.nodename_idx >= 0) ok &= (t.nodename_idx == p.node_base + i);
  // TaintToleration filter (taint_toleration.go:111-122) + raw score (taint_toleration.go:154-182)
  int raw = 0;
  {
    uint64_t untol = 0;
    #pragma unroll
    for (int w = 0; w < CCSIM_MAX_TAINT_WORDS; w++) {
      if (w < p.taint_words) {
        const uint64_t m = (w == 0) ? taint0 : p.taint_mask[(size_t)w * n + i];
        untol |= m & p.taint_nosched[w] & ~t.tol_nosched[w];
        raw += __popcll(m & p.taint_prefer[w] & ~t.tol_prefer[w]);
      }
    }
    if (t.filter_enable & CCSIM_PL_TAINT_TOLERATION) ok &= (untol == 0);
    if (!(t.score_enable & CCSIM_PL_TAINT_TOLERATION)) raw = 0;
  }
  // static-bit predicates: NodeAffinity (node_affinity.go:206-227), NodePorts (node_ports.go:157-185),
  // existing pods' anti-affinity (interpodaffinity/filtering.go:352-364)
  if (p.static_words > 0) {
    uint64_t sw[CCSIM_MAX_STATIC_WORDS];
    #pragma unroll
    for (int w = 0; w < CCSIM_MAX_STATIC_WORDS; w++) sw[w] = (w < p.static_words) ? p.static_mask[(size_t)w * n + i] : 0ull;
    if ((t.filter_enable & CCSIM_PL_NODE_AFFINITY) && (t.flags & (CCSIM_TF_HAS_NODE_SELECTOR | CCSIM_TF_HAS_AFFINITY_TERMS))) {
      bool m = true;
      #pragma unroll
      for (int w = 0; w < CCSIM_MAX_STATIC_WORDS; w++) m &= ((sw[w] & t.sel_mask[w]) == t.sel_mask[w]);
      if (t.flags & CCSIM_TF_HAS_AFFINITY_TERMS) {
        bool any = false;
        for (int k = 0; k < t.n_aff_terms; k++) {
          bool tm = true;
          #pragma unroll
          for (int w = 0; w < CCSIM_MAX_STATIC_WORDS; w++) tm &= ((sw[w] & t.aff_term_mask[k][w]) == t.aff_term_mask[k][w]);
          any |= tm;
        }
        m &= any;
      }
      ok &= m;
    }
    if ((t.filter_enable & CCSIM_PL_NODE_PORTS) && (t.flags & CCSIM_TF_HAS_HOST_PORTS)) {
      uint64_t c = 0;
      #pragma unroll
      for (int w = 0; w < CCSIM_MAX_STATIC_WORDS; w++) c |= sw[w] & t.port_static_mask[w];
      ok &= (c == 0);
    }
    if (t.filter_enable & CCSIM_PL_INTER_POD_AFFINITY) {
      uint64_t c = 0;
      #pragma unroll
      for (int w = 0; w < CCSIM_MAX_STATIC_WORDS; w++) c |= sw[w] & t.existing_anti_mask[w];
      ok &= (c == 0);
    }
  }
  if ((t.filter_enable & CCSIM_PL_NODE_PORTS) && (t.flags & CCSIM_TF_HAS_HOST_PORTS) && p.placed_mask)
    ok &= ((p.placed_mask[i] & t.port_tmpl_conflict) == 0);
  // NodeResourcesFit (fit.go:564-660)
  if (t.filter_enable & CCSIM_PL_FIT) {
    ok &= !(npods + 1 > a_pods);
    if (!(t.flags & CCSIM_TF_FIT_ALL_ZERO)) {
      ok &= !(t.req_cpu > 0 && t.req_cpu > a_cpu - r_cpu);
      ok &= !(t.req_mem > 0 && t.req_mem > a_mem - r_mem);
      if (t.req_eph > 0) ok &= !(t.req_eph > p.alloc_eph[i] - p.req_eph[i]);
      for (int k = 0; k < p.n_scalars; k++) {
        const int64_t q = t.req_scalar[k];
        if (q != 0) ok &= !(q > p.alloc_scalar[k][i] - p.req_scalar[k][i]);
      }
    }
  }
  // PodTopologySpread hard constraints (podtopologyspread/filtering.go:311-356)
  if (t.filter_enable & CCSIM_PL_POD_TOPOLOGY_SPREAD) {
    for (int c = 0; c < t.n_pts; c++) {
      const ccsim_pts &pc = t.pts[c];
      const DevCounter &dc = p.counters[pc.counter];
      const int32_t dom = dc.topo_col < 0 ? i : p.topo[dc.topo_col][i];
      if (dom < 0) { ok = false; break; }
      const long long skew = (long long)counter_base(p, pc.counter, smem_cnt)[dom] + pc.self_match - (long long)ptsmin[c];
      ok &= !(skew > pc.max_skew);
    }
  }
  // InterPodAffinity required terms (interpodaffinity/filtering.go:367-432)
  if (t.filter_enable & CCSIM_PL_INTER_POD_AFFINITY) {
    if (t.n_aff > 0) {
      bool pods_exist = true, missing = false;
      for (int a = 0; a < t.n_aff; a++) {
        const DevCounter &dc = p.counters[t.aff_counter[a]];
        const int32_t dom = dc.topo_col < 0 ? i : p.topo[dc.topo_col][i];
        if (dom < 0) { missing = true; break; }
        if (counter_base(p, t.aff_counter[a], smem_cnt)[dom] <= 0) pods_exist = false;
      }
      ok &= !(missing || (!pods_exist && !(aff_total == 0 && (t.flags & CCSIM_TF_AFF_SELF_MATCH_ALL))));
    }
    for (int a = 0; a < t.n_anti; a++) {
      const DevCounter &dc = p.counters[t.anti_counter[a]];
      const int32_t dom = dc.topo_col < 0 ? i : p.topo[dc.topo_col][i];
      if (dom >= 0) ok &= !(counter_base(p, t.anti_counter[a], smem_cnt)[dom] > 0);
    }
  }
  if (!ok) return false;
  // node-local score (framework.go:1137-1244: plugin score * weight, summed)
  int64_t sc = 0;
  if (t.score_enable & CCSIM_PL_FIT) sc += (int64_t)t.w_fit * score_least(a_cpu, a_mem, z_cpu, z_mem, t);
  if ((t.score_enable & CCSIM_PL_BALANCED) && !(t.flags & CCSIM_TF_BALANCED_SKIP))
    sc += (int64_t)t.w_balanced * score_balanced(a_cpu, a_mem, r_cpu, r_mem, t);
  cls = raw;
  score = sc;
  return true;
}

// TaintToleration NormalizeScore, reverse (helper/normalize_score.go:28-56)
__device__ __forceinline__ int64_t taint_norm(int raw, int maxraw) {
  if (maxraw == 0) return 100;
  return 100 - (100 * (int64_t)raw / maxraw);
}
