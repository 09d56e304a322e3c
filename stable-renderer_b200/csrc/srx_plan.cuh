// srx_plan.cuh — the overlap plan object shared by srx_overlap.cu (split kernels) and srx_fused.cu (persistent step).
#pragma once
#include "srx_common.cuh"

#define SRX_MAX_PEERS 8

struct srx_plan {
    srx_plan_desc d;
    bool fast_r8 = false;
    int64_t kcap = 0, n_valid = -1, key_min = 0, key_max = -1;
    // device tables (one allocation)
    int *tables = nullptr;
    int *colcell = nullptr, *rowcell = nullptr, *fmap = nullptr;
    // workspace (caller owned)
    char *ws = nullptr;
    int64_t ws_bytes = 0;
    int64_t accum_off = 0, accum_bytes = 0, winner_off = 0, winner_bytes = 0, winner64_off = 0, winner64_bytes = 0,
            status_off = 0, total_bytes = 0;
    int elem = 4;
    int cluster = 1;
    // frame-sharded peer mode (srx_plan_bind_peers): double-buffered accumulators + signal pads inside the workspace
    int64_t accum_stride = 0, pads_off = 0, ctrl_off = 0, stats_off = 0, stats_bytes = 0, ll_off = 0, ll_bytes = 0;
    int world = 1, rank = 0;
    char *peers[SRX_MAX_PEERS] = {nullptr};
    char *mc = nullptr;   // multicast (NVLS) address of the workspace, NULL = pull exchange
    int fused_grid = 0;   // CTAs of the persistent kernel; 0 = one per SM
    // cached plan (srx_plan_build_cache)
    int64_t need_off = 0, ctatab_off = 0, cntp_off = 0, cache_entries_cap = -1;
    void *pool = nullptr;
    int cache_grid = 0;
    bool cache_ready = false;
    bool fused = false;   // the persistent single-kernel step applies (fast_r8, float accumulators, aligned rows)
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static inline void plan_layout(srx_plan *p) {
    const srx_plan_desc &d = p->d;
    p->elem = d.accum_mode == SRX_ACCUM_DETERMINISTIC ? 8 : 4;
    int64_t off = 0;
    // [A0][A1][A2][signal pads][control words] come first: their offsets depend only on the key capacity and the
    // channel count, so they are identical on every rank of a frame-sharded run.
    //   A0/A1  accumulators of the persistent step kernel (srx_fused.cu), alternating with the device step counter;
    //          the kernel of step s clears the buffer of step s+1, so no memset node sits between steps
    //   A2     accumulator of the split reduce / gather entry points (cleared by a memset after each gather)
    p->accum_bytes = p->kcap * (d.channels + 1) * p->elem;
    p->accum_stride = align_up(p->accum_bytes, 256);
    off = 2 * p->accum_stride;
    p->accum_off = off;
    off += p->accum_stride;
    p->pads_off = off;   // [3 barriers][SRX_MAX_PEERS sources] u32 arrival counters (monotonic)
    off += 256;
    p->ctrl_off = off;   // [0] step counter; +64 phase stamps of the first / last CTA; +256 ring of per-step (first CTA
    off += 4096;         // start, last CTA end) global-timer stamps; +1024 ring of the first CTA's phase stamps (profiling aids)
    // peer mode: exchanged totals as 32-byte records.  Pull exchange: this rank's slice, [ceil(K / world)] records
    // {sum.xyzw, count, step}, world >= 2.  NVLS exchange: all K records (every owner broadcasts its slice); tables of more
    // than 4 Mi slots keep the half-size region and the pull exchange.
    p->ll_off = off;
    p->ll_bytes = p->fused ? (p->kcap <= (4ll << 20) ? p->kcap + 64 : p->kcap / 2 + 64) * 32 : 0;
    off = align_up(off + p->ll_bytes, 256);
    p->winner_off = off;
    p->winner_bytes = (int64_t)d.batch * d.lat_h * d.lat_w * 4;
    off = align_up(off + p->winner_bytes, 256);
    p->winner64_off = off;
    p->winner64_bytes = p->fast_r8 ? 0 : (int64_t)d.batch * d.lat_h * d.lat_w * 8;
    off = align_up(off + p->winner64_bytes, 256);
    p->status_off = off;
    off += 256;
    p->stats_off = off;  // [<= 1024 CTAs + batch][6] statistics records of 32 B {3 doubles, step}: the AdaIN partial sums of
                         // (CTA c, frame f) live in slot c + f; exchanged between the CTAs that share a latent frame
    p->stats_bytes = ((int64_t)1024 + d.batch) * 6 * 32;
    off = align_up(off + p->stats_bytes, 256);
    p->need_off = off;   // [K] byte map of winner keys (cached plan); bytes so that ranks can combine theirs with a MAX all-reduce
    off = align_up(off + p->kcap, 256);
    p->cntp_off = off;   // [K] f32 entries per key of this rank's ids (cached plan)
    off = align_up(off + (p->fused ? p->kcap * 4 : 0), 256);
    p->ctatab_off = off; // [3][<= 1024 CTAs] per-CTA pair counts / kept entries / region starts (cached plan)
    off += 3 * 1024 * 4;
    p->total_bytes = off;
}


// srx_fused.cu
bool srx_fused_applicable(const srx_plan *p);
int srx_launch_fused(srx_plan *p, const srx_step_args *a, cudaStream_t st);
