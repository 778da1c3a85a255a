// lsm_persist.cuh -- the whole Longstaff-Schwartz backward induction (LSMPricer.cpp:35-101) in ONE cooperative launch.
// Included by lsm.cu after the packed-fp32 step arithmetic (fast2_compute) and the bulk-copy helpers.
//
// Why.  One launch per time step costs, per step, a kernel boundary plus a serial tail: per-CTA partials -> ticket ->
// last CTA folds -> (multi-GPU) exchange -> one thread solves -> only then may the dependent launch proceed.  At 2^23
// paths per GPU (config 3 on 8 GPUs) that tail is a third of the step.  Here the grid stays resident for all M steps:
//
//   worker CTAs (blockIdx 1 .. nw)   own a FIXED set of 4096-path tiles for the whole induction and stream them through a
//       TMA ring whose slab copies never drain: the tile sequence (step s, tile it) is one continuous stream, the slab
//       tiles of step s+1 are in flight while step s is still being computed, and because a worker's carry tiles are
//       written by that worker only, their copies need no grid barrier -- one proxy fence per step (generic-proxy
//       stores -> fence.proxy.async -> bulk copy), executed while the worker waits for the coefficients anyway.  Per step
//       a worker sends ONE row of regression moments to the reducer and waits for ONE row of coefficients.
//   reducer CTA (blockIdx 0)         does no streaming.  It polls the workers' rows, folds them in worker order, pushes
//       the folded row straight into every peer GPU's mailbox (plain P2P stores over NVLink, 8-byte words that carry
//       32 bits of payload and a 32-bit sequence tag, so delivery and publication are one store and one hop), polls its
//       own mailbox, adds the rows in rank order (every rank obtains the bitwise identical sum), solves the
//       (p+1)x(p+1) system and broadcasts the coefficients to its workers.  No NCCL call, no host round trip and no
//       kernel boundary anywhere in the induction; the standardisation sample sums, the final {sum V0, N} and the
//       sum of squared deviations for the standard error go through the same exchange.
//
// All waits are bounded (PX_TIMEOUT_NS): a rank that never arrives raises the error flag, the reducer tells its
// workers and its peers to stop, and the host call fails with MCP_ERR_NCCL instead of hanging or pricing partial sums.
#pragma once

namespace px {

constexpr int NT = 512;                       // threads per CTA (workers: 8 paths per thread and tile)
constexpr int TILE = 4096;                    // paths per tile
constexpr int STAGE_FLOATS = 3 * TILE;        // S_j | S_{j-1} | V
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
constexpr unsigned long long PX_TIMEOUT_NS = 20000000000ull;
constexpr int PH0_ROWS = 256;                 // standardisation rows (3 doubles each) per exchange
constexpr int RED_GOT_BYTES = 104 * 1024;     // reducer: gather area (<= 16 ranks x (2 * 3 * PH0_ROWS + 1) words)
constexpr int RED_VALS = 1024;                // reducer: doubles staged for an exchange; fold scratch [NSEG][NV] behind them
constexpr int RED_SMEM_BYTES = RED_GOT_BYTES + (RED_VALS + 512) * 8;
static_assert(MCP_XMAX_RANKS * (6 * PH0_ROWS + 1) * 4 <= RED_GOT_BYTES, "gather area");
static_assert(6 * PH0_ROWS + 1 <= MCP_PX_BIGW && 3 * PH0_ROWS <= RED_VALS, "phase-0 chunk");
static_assert(MCP_PX_MAXW * 2 * 20 * 4 <= RED_GOT_BYTES, "worker rows");

struct Args {
    const float* S;
    int64_t ld, n;
    float* V;
    int32_t* tau;
    double* coef;      // [M][COEF_LD]  (output table; the reducer fills row j-1 at step j)
    double* mu;        // [M]           standardisation, produced by the reducer in phase 0
    double* inv_s;     // [M]
    double* ssum;      // [M][4]        local sample sums (workers -> reducer)
    double* fin;       // [4]           sum V0, sum (V0-mean)^2, N (all ranks)
    const int* kind;   // [M]
    double K, disc;
    StepK g;           // contract constants of the step arithmetic (host-made: they live in the constant bank)
    int is_call, M, ns, l2_resident, n_workers, n_stages, ref_rank;
    McpPx x;
    unsigned long long seq0;  // tag of this launch's first exchange; exchange e uses seq0 + e
    unsigned long long* trace;  // optional [M][n_workers + 1][4] globaltimer stamps (tools/px_trace.py); nullptr = off
};

__device__ __forceinline__ unsigned long long tag_of(unsigned long long seq) { return ((seq % 0xffffffffull) + 1ull) << 32; }
__device__ __forceinline__ unsigned long long ld_word(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_word(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long* local_row(const McpPx& x, unsigned long long seq, int w) {
    return x.local + ((size_t)(seq & 1ull) * MCP_PX_MAXW + (size_t)w) * MCP_PX_ROWW;
}
__device__ __forceinline__ unsigned long long* local_bc(const McpPx& x, unsigned long long seq, int w) {
    return x.local + (size_t)2 * MCP_PX_MAXW * MCP_PX_ROWW + ((size_t)(seq & 1ull) * MCP_PX_MAXW + (size_t)w) * MCP_PX_BCW;
}
__device__ __forceinline__ unsigned long long* peer_slot(const McpPx& x, int dst_rank, unsigned long long seq, int src_rank) {
    return x.peer[dst_rank] + MCP_XLEGACY_WORDS + ((size_t)(seq & 1ull) * MCP_XMAX_RANKS + (size_t)src_rank) * MCP_PX_BIGW;
}
// double k of a tagged row = words 2k (low half) and 2k+1 (high half)
__device__ __forceinline__ void st_tagged_double(unsigned long long* row, int k, double v, unsigned long long tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    st_word(row + 2 * k, (b & 0xffffffffull) | tag);
    st_word(row + 2 * k + 1, (b >> 32) | tag);
}
// spin until `*p` carries `tag`; false after PX_TIMEOUT_NS
__device__ __forceinline__ bool wait_word(const unsigned long long* p, unsigned long long tag, unsigned int* lo) {
    unsigned long long w = ld_word(p);
    if ((w & 0xffffffff00000000ull) != tag) {
        const unsigned long long t0 = global_ns();
        unsigned int spins = 0;
        do {
            w = ld_word(p);
            if ((w & 0xffffffff00000000ull) == tag) break;
            if ((++spins & 1023u) == 0u && global_ns() - t0 > PX_TIMEOUT_NS) return false;  // the timer is read once per 1024 polls
        } while (true);
    }
    *lo = (unsigned int)w;
    return true;
}

// ------------------------------------------------------------------------------------------------ reducer side
// Gather `nwords` tagged words from each of `nrows` rows (row r at base + r * stride) into got[r * nwords + k].
// All NT threads poll in parallel, several independent loads in flight per thread.
__device__ __forceinline__ bool gather_rows(const unsigned long long* base, size_t stride, int nrows, int nwords, unsigned long long tag,
                                            unsigned int* got, int* s_fail) {
    const int total = nrows * nwords;
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < total; i0 += NT * U) {
        const unsigned long long* p[U];
        unsigned long long w[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * NT;
            live[u] = i < total;
            const int r = live[u] ? i / nwords : 0, k = live[u] ? i - r * nwords : 0;
            p[u] = base + (size_t)r * stride + k;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) w[u] = live[u] ? ld_word(p[u]) : tag;
        bool all = true;
#pragma unroll
        for (int u = 0; u < U; ++u) all = all && ((w[u] & 0xffffffff00000000ull) == tag);
        if (!all) {
            const unsigned long long t0 = global_ns();
            unsigned int spins = 0;
            while (true) {
                all = true;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((w[u] & 0xffffffff00000000ull) != tag) w[u] = ld_word(p[u]);
                    all = all && ((w[u] & 0xffffffff00000000ull) == tag);
                }
                if (all) break;
                if ((++spins & 255u) == 0u && (global_ns() - t0 > PX_TIMEOUT_NS || *(volatile int*)s_fail)) { *(volatile int*)s_fail = 1; break; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (live[u]) got[i0 + u * NT] = (unsigned int)w[u];
    }
    __syncthreads();
    return *(volatile int*)s_fail == 0;
}

__device__ __forceinline__ double got_double(const unsigned int* got, int row, int nwords, int k) {
    const unsigned int* g = got + (size_t)row * nwords + 2 * k;
    return __longlong_as_double((long long)(((unsigned long long)g[1] << 32) | (unsigned long long)g[0]));
}

// vals[0..nv) (shared) <- sum over ranks, in rank order, of every rank's vals.  One hop: each rank stores its row (plus a
// status word) into every rank's slot.  Returns false on time-out or when a peer reports that it is giving up.
__device__ __forceinline__ bool exchange(const McpPx& x, unsigned long long seq, double* vals, int nv, unsigned int* got, int* s_fail, int my_status) {
    if (x.nranks <= 1) return my_status == 0;
    const unsigned long long tag = tag_of(seq);
    const int nwords = 2 * nv + 1;
    MCP_DBG_CHECK(nwords <= MCP_PX_BIGW && x.nranks <= MCP_XMAX_RANKS && x.rank >= 0 && x.rank < x.nranks, DBG_XCHG_ROW);
    for (int i = threadIdx.x; i < x.nranks * nwords; i += NT) {
        const int r = i / nwords, k = i - r * nwords;
        unsigned long long payload;
        if (k == 2 * nv) payload = (unsigned long long)(unsigned int)my_status;
        else {
            const unsigned long long b = (unsigned long long)__double_as_longlong(vals[k >> 1]);
            payload = (k & 1) ? (b >> 32) : (b & 0xffffffffull);
        }
        st_word(peer_slot(x, r, seq, x.rank) + k, payload | tag);
    }
    const bool ok = gather_rows(peer_slot(x, x.rank, seq, 0), MCP_PX_BIGW, x.nranks, nwords, tag, got, s_fail);
    if (!ok) return false;
    bool peers_ok = my_status == 0;
    for (int r = 0; r < x.nranks; ++r) peers_ok = peers_ok && got[(size_t)r * nwords + 2 * nv] == 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < nv; k += NT) {
        double s = 0.0;
        for (int r = 0; r < x.nranks; ++r) s += got_double(got, r, nwords, k);
        vals[k] = s;
    }
    __syncthreads();
    return peers_ok;
}

// Reducer -> every worker's own slot: word 0 = status (0 = go on, 1 = stop), words 1 .. 2 nv = the doubles as (low, high) halves.
static_assert(2 * (MAXP + 1) + 1 <= MCP_PX_BCW, "broadcast slot");
__device__ __forceinline__ void broadcast(const McpPx& x, unsigned long long seq, int nw, const double* vals, int nv, int status) {
    const unsigned long long tag = tag_of(seq);
    const int nwords = 2 * nv + 1;
    MCP_DBG_CHECK(nwords <= MCP_PX_BCW && nw >= 1 && nw <= MCP_PX_MAXW, DBG_XCHG_ROW);
    for (int i = threadIdx.x; i < nw * nwords; i += NT) {
        const int w = i / nwords, k = i - w * nwords;
        unsigned long long payload = (unsigned long long)(unsigned int)status;
        if (k > 0) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(vals[(k - 1) >> 1]);
            payload = ((k - 1) & 1) ? (b >> 32) : (b & 0xffffffffull);
        }
        st_word(local_bc(x, seq, w) + k, payload | tag);
    }
}

template <int P>
__device__ void reducer(const Args& a, unsigned int* got /* shared, >= max(nw * 2NV, nranks * (6M'+1)) words */, double* vals /* shared [>= 64] */) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    __shared__ int s_fail;
    __shared__ double s_coef[COEF_LD];
    const int tid = threadIdx.x, nw = a.n_workers, M = a.M;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    unsigned long long seq = a.seq0;
    bool ok = true;
    auto give_up = [&](unsigned long long at_seq) {  // tell the workers (and, through the next push, nobody: peers time out or see status)
        if (tid == 0) *a.x.err = 1;
        broadcast(a.x, at_seq, nw, vals, 0, 1);
    };

    // ---- phase 0: standardisation tables from the sample sums of every rank -------------------------------------
    // workers left ssum[j][0..2] in global memory and raised a tagged flag in their row
    ok = gather_rows(local_row(a.x, seq, 0), MCP_PX_ROWW, nw, 1, tag_of(seq), got, &s_fail);
    __threadfence();
    {
        const unsigned long long seq_bc = seq;  // the workers wait on the phase's first sequence number
        for (int j0 = 0; j0 < M; j0 += PH0_ROWS) {
            const int nn = M - j0 < PH0_ROWS ? M - j0 : PH0_ROWS;
            __syncthreads();
            for (int i = tid; i < nn * 3; i += NT) vals[i] = __ldcg(a.ssum + (size_t)(j0 + i / 3) * 4 + (i % 3));
            __syncthreads();
            ok = exchange(a.x, seq, vals, nn * 3, got, &s_fail, ok ? 0 : 1) && ok;
            ++seq;
            for (int i = tid; i < nn; i += NT) {  // lsm_scale_finalize: mean / std of the in-the-money sample
                const int j = j0 + i;
                const double cnt = vals[3 * i], s1 = vals[3 * i + 1], s2 = vals[3 * i + 2];
                double m = a.K, sd = fabs(a.K) > 0.0 ? fabs(a.K) : 1.0;
                if (cnt >= 2.0) {
                    m = s1 / cnt;
                    const double var = (s2 - cnt * m * m) / (cnt - 1.0);
                    if (var > 1e-12 * m * m) sd = sqrt(var);
                    else sd = fabs(m) > 0.0 ? fabs(m) : 1.0;
                }
                a.mu[j] = m;
                a.inv_s[j] = 1.0 / sd;
            }
        }
        __threadfence();
        __syncthreads();
        if (!ok) { give_up(seq_bc); return; }
        broadcast(a.x, seq_bc, nw, vals, 0, 0);
    }

    // ---- the induction: one exchange per step that regresses, one for {sum V0, N}, one for the squared deviations ----
    for (int j = M - 1; j >= 0; --j) {
        const bool dm = j > 0 && a.kind[j - 1] == 0, df = j == 0;
        if (!dm && !df) continue;
        const unsigned long long tag = tag_of(seq);
        ok = gather_rows(local_row(a.x, seq, 0), MCP_PX_ROWW, nw, 2 * NV, tag, got, &s_fail);
        if (a.trace && tid == 0) a.trace[((size_t)(M - 1 - j) * (nw + 1)) * 4 + 0] = global_ns();  // all rows in
        // fold the workers' rows in worker order: thread (seg, k) adds a contiguous block of rows, thread k the block sums
        constexpr int NSEG = NT / NV;
        double* seg_sum = vals + RED_VALS;  // [NSEG][NV]
        const int rows = (nw + NSEG - 1) / NSEG;
        if (tid < NSEG * NV) {
            const int seg = tid / NV, k = tid - seg * NV;
            const int b0 = seg * rows, b1 = min(b0 + rows, nw);
            double s = 0.0;
            for (int b = b0; b < b1; ++b) s += got_double(got, b, 2 * NV, k);
            seg_sum[seg * NV + k] = s;
        }
        __syncthreads();
        if (tid < NV) {
            double t = 0.0;
#pragma unroll
            for (int g = 0; g < NSEG; ++g) t += seg_sum[g * NV + tid];
            vals[tid] = t;
        }
        if (df && tid == 1) vals[1] = (double)a.n;  // the final exchange carries {sum V0, N}
        __syncthreads();
        if (a.trace && tid == 0) a.trace[((size_t)(M - 1 - j) * (nw + 1)) * 4 + 1] = global_ns();  // folded
        ok = exchange(a.x, seq, vals, NV, got, &s_fail, ok ? 0 : 1) && ok;
        if (a.trace && tid == 0) a.trace[((size_t)(M - 1 - j) * (nw + 1)) * 4 + 2] = global_ns();  // exchanged
        if (!ok) { give_up(seq); return; }
        if (df) {
            if (tid == 0) {
                a.fin[0] = vals[0];
                a.fin[2] = vals[1];
                s_coef[0] = vals[0] / vals[1];  // the mean: workers need it for the squared deviations
            }
            __syncthreads();
            broadcast(a.x, seq, nw, s_coef, 1, 0);
        } else {
            if (tid == 0) {
                const RefRank rr{__ldcg(a.mu + j - 1), __ldcg(a.inv_s + j - 1)};
                solve_normal_equations<P>(vals, s_coef, a.ref_rank ? &rr : nullptr);
            }
            __syncthreads();
            if (tid < COEF_LD) a.coef[(size_t)(j - 1) * COEF_LD + tid] = s_coef[tid];
            broadcast(a.x, seq, nw, s_coef, P + 1, 0);
        }
        if (a.trace && tid == 0) a.trace[((size_t)(M - 1 - j) * (nw + 1)) * 4 + 3] = global_ns();  // solved and broadcast
        ++seq;
    }
    // ---- sum of squared deviations ----
    {
        ok = gather_rows(local_row(a.x, seq, 0), MCP_PX_ROWW, nw, 2, tag_of(seq), got, &s_fail);
        if (tid == 0) {
            double t = 0.0;
            for (int b = 0; b < nw; ++b) t += got_double(got, b, 2, 0);
            vals[0] = t;
        }
        __syncthreads();
        ok = exchange(a.x, seq, vals, 1, got, &s_fail, ok ? 0 : 1) && ok;
        if (tid == 0) {
            a.fin[1] = vals[0];
            if (!ok) *a.x.err = 1;
        }
    }
}

// number of exchanges (sequence numbers) one launch consumes -- the host advances ctx->xchg_seq by this, identically on every rank
inline unsigned long long exchanges_per_launch(int M, const int* kind) {
    unsigned long long e = (unsigned long long)((M + PH0_ROWS - 1) / PH0_ROWS);
    for (int j = M - 1; j >= 0; --j)
        if ((j > 0 && kind[j - 1] == 0) || j == 0) ++e;
    return e + 1;
}

// ------------------------------------------------------------------------------------------------- worker side
template <int P, bool TAU>
__device__ void worker(const Args& a, unsigned char* smem_raw) {
    constexpr int NM = 3 * P + 2;
    constexpr int NV = NM > 2 ? NM : 2;
    constexpr int FLUSH = 8;
    const int n_stages = a.n_stages;
    float* ring = reinterpret_cast<float*>(smem_raw);
    unsigned char* after_ring = smem_raw + (size_t)n_stages * STAGE_BYTES + MCP_DBG_CANARY_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(after_ring);
    uint64_t* empty = full + 8;
    double* sacc = reinterpret_cast<double*>(after_ring + 128);  // [NV][NT / 2]: lanes 2k and 2k+1 share a slot (owned by the even lane),
                                                                 // which buys the ring a fourth stage
#ifdef MCP_DEBUG_BOUNDS
    if (threadIdx.x < MCP_DBG_CANARY_BYTES / 4) reinterpret_cast<unsigned int*>(after_ring - MCP_DBG_CANARY_BYTES)[threadIdx.x] = MCP_DBG_CANARY;
#endif
    __shared__ double red[NT / 32][NV];
    __shared__ double s_c[COEF_LD + 4];  // c_0..c_P, then mu_j, 1/s_j, mu_{j-1}, 1/s_{j-1}
    __shared__ int s_stop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = (int)blockIdx.x - 1, nw = a.n_workers, M = a.M;
    const float* __restrict__ S = a.S;
    float* __restrict__ V = a.V;

    const int64_t ntile = (a.n + TILE - 1) / TILE;
    const int64_t my_tiles = w < ntile ? (ntile - 1 - w) / nw + 1 : 0;  // tiles w, w + nw, ...
    const int64_t g_total = my_tiles * M;                               // the whole stream of this worker
    // The ring runs ahead across step boundaries when the carry tile it fetches was written at least one full ring earlier in
    // the stream; a worker with fewer tiles than that refills once per step instead (tiny problems only).
    const bool chain = my_tiles > n_stages;
    // only the last tile of the path set can be ragged, and it is the last tile of the worker that owns it
    const int n_whole = (int)my_tiles - ((my_tiles > 0 && ((int64_t)w + (my_tiles - 1) * nw + 1) * TILE > a.n) ? 1 : 0);
    const uint64_t pol_first = l2_policy_evict_first(), pol_norm = l2_policy_evict_normal();
    unsigned long long seq = a.seq0;

    // The stream of this worker = (step 0, tiles 0 .. my_tiles-1), (step 1, ...), ...; stream tile g uses ring stage
    // g mod n_stages.  One elected thread walks it with two cursors (no divisions in the loop): the SLAB cursor arms a stage's
    // barrier with the bytes of all its copies and starts the copies of S_j | S_{j-1}; the CARRY cursor follows it and starts
    // the copy of V.  The slab cursor runs ahead across step boundaries; the carry cursor enters a step only after that step's
    // top-of-loop proxy fence.
    struct Cursor { int s; int64_t it; int st; int64_t n; };
    Cursor cs = {0, 0, 0, 0}, cv = {0, 0, 0, 0};  // thread 0 only
    auto advance = [&](Cursor& c) {
        ++c.n;
        if (++c.it == my_tiles) { c.it = 0; ++c.s; }
        if (++c.st == n_stages) c.st = 0;
    };
    auto tile_bytes = [&](int64_t i0) -> uint32_t {
        const int64_t cnt = a.ld - i0 < TILE ? a.ld - i0 : TILE;  // rows are padded to ld (multiple of 128)
        return (uint32_t)cnt * 4u;
    };
    auto issue_slab = [&]() {
        const int j = M - 1 - cs.s;
        const bool dm = j > 0 && __ldg(a.kind + j - 1) == 0, want_v = cs.s > 0;
        const int64_t i0 = ((int64_t)w + cs.it * nw) * TILE;
        const uint32_t bytes = tile_bytes(i0);
        float* dst = ring + (size_t)cs.st * STAGE_FLOATS;
        MCP_DBG_CHECK(cs.st >= 0 && cs.st < n_stages && cs.it >= 0 && cs.it < my_tiles && cs.s >= 0 && cs.s < M && cs.n < g_total, DBG_RING_STAGE);
        MCP_DBG_CHECK(i0 >= 0 && bytes > 0 && bytes <= TILE * 4u && i0 + (int64_t)(bytes / 4u) <= a.ld && (bytes & 15u) == 0u && j >= 0 && j < M, DBG_RING_ISSUE);
        mbar_expect_tx(full + cs.st, bytes * (1u + (dm ? 1u : 0u) + (want_v ? 1u : 0u)));
        bulk_g2s_hint(dst, S + (int64_t)j * a.ld + i0, bytes, full + cs.st, pol_first);                      // last use of row j
        if (dm) bulk_g2s_hint(dst + TILE, S + (int64_t)(j - 1) * a.ld + i0, bytes, full + cs.st, pol_norm);  // read again next step
        advance(cs);
    };
    auto issue_carry = [&]() {
        MCP_DBG_CHECK(cv.n < cs.n && cv.st >= 0 && cv.st < n_stages, DBG_RING_STAGE);  // the carry cursor never passes the slab cursor
        if (cv.s > 0) {
            const int64_t i0 = ((int64_t)w + cv.it * nw) * TILE;
            bulk_g2s_hint(ring + (size_t)cv.st * STAGE_FLOATS + 2 * TILE, V + i0, tile_bytes(i0), full + cv.st, a.l2_resident ? pol_norm : pol_first);
        }
        advance(cv);
    };
    if (tid == 0) {
        s_stop = 0;
        for (int st = 0; st < n_stages; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the terminal step reads no carry: its tiles can fly before anything else happens
        while (cs.n < n_stages && cs.n < my_tiles) { issue_slab(); issue_carry(); }
    }
    for (int m = 0; m < NV; ++m)
        if (!(tid & 1)) sacc[m * (NT / 2) + (tid >> 1)] = 0.0;

    // ---- phase 0: sample sums of rows w, w + nw, ... over the first ns paths (lsm_scale_sums) ----
    for (int j = w; j < M; j += nw) {
        double acc[3] = {0.0, 0.0, 0.0};
        const float* row = S + (int64_t)j * a.ld;
        for (int i = tid; i < a.ns; i += NT) {
            const double s = (double)__ldg(row + i);
            if (payoff_fn(a.is_call, s, a.K) > 1e-14) { acc[0] += 1.0; acc[1] += s; acc[2] = fma(s, s, acc[2]); }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double s = warp_sum(acc[k]);
            if (lane == 0) red[warp][k] = s;
        }
        __syncthreads();
        if (tid < 3) {
            double s = 0.0;
            for (int q = 0; q < NT / 32; ++q) s += red[q][tid];
            __stcg(a.ssum + (size_t)j * 4 + tid, s);
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_word(local_row(a.x, seq, w), tag_of(seq));
    // wait for the tables
    const unsigned long long seq_tables = seq;
    seq += (unsigned long long)((M + PH0_ROWS - 1) / PH0_ROWS);
    if (tid == 0) {
        unsigned int status = 1;
        if (!wait_word(local_bc(a.x, seq_tables, w), tag_of(seq_tables), &status) || status != 0) s_stop = 1;
    }
    __syncthreads();
    __threadfence();

    FastConsts<P> k;
    SweepArgs sa;  // the view fast2_compute expects
    memset(&sa, 0, sizeof(sa));
    sa.n = a.n; sa.tau = a.tau;
    float2 la[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) la[m] = make_float2(0.f, 0.f);
    int since = 0, cnt = 0;
    int64_t g = 0;        // stream tiles consumed
    int cst = 0;          // ring stage / phase parity of the next tile to consume
    uint32_t cpar = 0;
    unsigned long long seq_coef = 0;  // sequence number whose broadcast carries the coefficients of the coming step
    double mean = 0.0;

    for (int s = 0; s < M && !s_stop; ++s) {
        const int j = M - 1 - s;
        const int mode = s == 0 ? 2 : __ldg(a.kind + j);
        const bool dm = j > 0 && __ldg(a.kind + j - 1) == 0, df = j == 0;
        // every carry store of the previous step is ordered before the bulk copies (async proxy) that fetch it back
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        if (tid == 0 && s > 0) {
            if (!chain)
                while (cs.s <= s && cs.n < g_total) issue_slab();
            while (cv.n < cs.n && cv.s <= s) issue_carry();  // carry parts of the tiles whose slab part ran ahead
        }
        // ---- constants of the step: c_j from the reducer's broadcast (the only wait of the step), mu / 1/s from the tables ----
        if (tid < 4) {
            const int jj = (tid < 2) ? j : (j > 0 ? j - 1 : 0);
            s_c[COEF_LD + tid] = __ldcg(((tid & 1) ? a.inv_s : a.mu) + jj);
        }
        if (mode == 0) {
            if (tid <= 2 * (P + 1)) {  // word 0 = status, words 1.. = c_0 .. c_P as (low, high) halves
                unsigned int lo = 1;
                const bool got_it = wait_word(local_bc(a.x, seq_coef, w) + tid, tag_of(seq_coef), &lo);
                if (tid == 0) { if (!got_it || lo != 0) s_stop = 1; }
                else reinterpret_cast<unsigned int*>(s_c)[tid - 1] = got_it ? lo : 0u;
            }
        }
        __syncthreads();
        if (s_stop) break;
        if (a.trace && tid == 0) a.trace[((size_t)s * (nw + 1) + (w + 1)) * 4 + 0] = global_ns();  // constants in hand
#pragma unroll
        for (int m = 0; m <= P; ++m) k.c[m] = splat2(mode == 0 ? (float)s_c[m] : 0.f);
        k.is = splat2((float)s_c[COEF_LD + 1]);
        k.c0 = splat2((float)(-s_c[COEF_LD] * s_c[COEF_LD + 1]));
        k.is_p = splat2((float)s_c[COEF_LD + 3]);
        k.c0_p = splat2((float)(-s_c[COEF_LD + 2] * s_c[COEF_LD + 3]));
        sa.j = j; sa.terminal = s == 0; sa.do_moments = dm; sa.do_final = df;

        // One tile of the stream.  WHOLE tiles (all but possibly the very last tile of the path set) take the path without
        // bounds tests; per-thread addresses advance by a constant stride.
        auto one_tile = [&](auto kind_tag, auto tail_tag, float* vp, int64_t i0) {
            constexpr int KIND = decltype(kind_tag)::value;
            constexpr bool TAILT = decltype(tail_tag)::value;
            const bool ldm = KIND == 0 ? true : dm, wv = KIND == 0 ? true : (mode != 2);
            const int st = cst;
            const uint32_t parity = cpar;
            while (!mbar_try_wait(full + st, parity)) {}
            const float* buf = ring + (size_t)st * STAGE_FLOATS + 4 * tid;
            F8 s8, p8, v8;
            {
                const float4 x0 = *reinterpret_cast<const float4*>(buf), x1 = *reinterpret_cast<const float4*>(buf + 2048);
                s8.q[0] = make_float2(x0.x, x0.y); s8.q[1] = make_float2(x0.z, x0.w); s8.q[2] = make_float2(x1.x, x1.y); s8.q[3] = make_float2(x1.z, x1.w);
            }
            p8 = s8; v8 = s8;
            if (ldm) {
                const float4 x0 = *reinterpret_cast<const float4*>(buf + TILE), x1 = *reinterpret_cast<const float4*>(buf + TILE + 2048);
                p8.q[0] = make_float2(x0.x, x0.y); p8.q[1] = make_float2(x0.z, x0.w); p8.q[2] = make_float2(x1.x, x1.y); p8.q[3] = make_float2(x1.z, x1.w);
            }
            if (wv) {
                const float4 x0 = *reinterpret_cast<const float4*>(buf + 2 * TILE), x1 = *reinterpret_cast<const float4*>(buf + 2 * TILE + 2048);
                v8.q[0] = make_float2(x0.x, x0.y); v8.q[1] = make_float2(x0.z, x0.w); v8.q[2] = make_float2(x1.x, x1.y); v8.q[3] = make_float2(x1.z, x1.w);
            }
            // this warp holds its part of the stage in registers; once all 16 warps have said so the slot is refilled: with
            // the next tile of this step (slab and carry), or with the slab part of a tile of the next step
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + st);
            if (tid == 0 && chain && cs.n < g_total) {
                while (!mbar_try_wait(empty + st, parity)) {}
                issue_slab();
                while (cv.n < cs.n && cv.s <= s) issue_carry();
            }
            if (++cst == n_stages) { cst = 0; cpar ^= 1u; }
            const int64_t ia = i0 + 4 * tid, ib = ia + 2048;
            fast2_compute<P, TAU, TAILT, KIND>(sa, a.g, k, s8, p8, v8, ia, ib, mode, la, cnt);
            MCP_DBG_CHECK(vp == V + ia && ia >= 0 && (!(!TAILT || ia < a.ld) || ia + 4 <= a.ld) && (!(!TAILT || ib < a.ld) || ib + 4 <= a.ld), DBG_CARRY_STORE);
            MCP_DBG_CHECK(cst >= 0 && cst < n_stages, DBG_RING_STAGE);
            if (!TAILT || ia < a.ld) stg4_keep(vp, v8.q[0], v8.q[1]);
            if (!TAILT || ib < a.ld) stg4_keep(vp + 2048, v8.q[2], v8.q[3]);
            if (++since == FLUSH) {
#pragma unroll
                for (int m = 0; m < NV; ++m) {
                    float v = la[m].x + la[m].y;
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if (!(tid & 1)) sacc[m * (NT / 2) + (tid >> 1)] += (double)v;
                    la[m] = make_float2(0.f, 0.f);
                }
                since = 0;
            }
        };
        auto run_tiles = [&](auto kind_tag) {
            float* vp = V + (int64_t)w * TILE + 4 * tid;
            const int64_t vstride = (int64_t)nw * TILE;
            int64_t i0 = (int64_t)w * TILE;  // (dead code in the whole-tile loop unless first-exercise indices are wanted)
            for (int it = 0; it < n_whole; ++it, vp += vstride, i0 += vstride) one_tile(kind_tag, std::false_type{}, vp, i0);
            if (n_whole < (int)my_tiles) one_tile(kind_tag, std::true_type{}, vp, i0);
            g += my_tiles;
        };
        if (mode == 0 && dm && !df) run_tiles(std::integral_constant<int, 0>{});
        else run_tiles(std::integral_constant<int, 1>{});

        if (a.trace && tid == 0) a.trace[((size_t)s * (nw + 1) + (w + 1)) * 4 + 1] = global_ns();      // thread 0 through its tiles
        if (dm || df) {
            // ---- this worker's row of the step: warp sums -> 16 rows in shared memory -> thread m adds them in warp order ----
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                const double mine = (tid & 1) ? 0.0 : sacc[m * (NT / 2) + (tid >> 1)];
                const double t = warp_sum(mine + (double)(la[m].x + la[m].y) + (m == 0 ? (double)cnt : 0.0));
                if (lane == 0) red[warp][m] = t;
                if (!(tid & 1)) sacc[m * (NT / 2) + (tid >> 1)] = 0.0;
                la[m] = make_float2(0.f, 0.f);
            }
            since = 0;
            cnt = 0;
            __syncthreads();
            MCP_DBG_CHECK(w >= 0 && w < MCP_PX_MAXW && 2 * NV <= MCP_PX_ROWW, DBG_XCHG_ROW);
            if (tid < NV) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < NT / 32; ++q) t += red[q][tid];
                st_tagged_double(local_row(a.x, seq, w), tid, t, tag_of(seq));
            }
            if (a.trace && tid == 0) a.trace[((size_t)s * (nw + 1) + (w + 1)) * 4 + 2] = global_ns();  // row sent
            seq_coef = seq;
            ++seq;
        }
    }

    // ---- standard error: the reducer's last broadcast carried the global mean ----
    if (!s_stop) {
        if (tid < 3) {
            unsigned int lo = 1;
            const bool got_it = wait_word(local_bc(a.x, seq_coef, w) + tid, tag_of(seq_coef), &lo);
            if (tid == 0) { if (!got_it || lo != 0) s_stop = 1; }
            else reinterpret_cast<unsigned int*>(s_c)[tid - 1] = got_it ? lo : 0u;
        }
        __syncthreads();
        mean = s_c[0];
    }
    if (!s_stop) {
        double acc = 0.0;
        for (int64_t it = 0; it < my_tiles; ++it) {
            const int64_t i0 = ((int64_t)w + it * nw) * TILE;
            for (int e = tid; e < TILE; e += NT) {
                const int64_t i = i0 + e;
                if (i < a.n) {
                    const double d = (double)__ldcg(V + i) - mean;
                    acc = fma(d, d, acc);
                }
            }
        }
        const double t = warp_sum(acc);
        __syncthreads();
        if (lane == 0) red[warp][0] = t;
        __syncthreads();
        if (tid == 0) {
            double tt = 0.0;
            for (int q = 0; q < NT / 32; ++q) tt += red[q][0];
            st_tagged_double(local_row(a.x, seq, w), 0, tt, tag_of(seq));
        }
    }
#ifdef MCP_DEBUG_BOUNDS
    if (tid < MCP_DBG_CANARY_BYTES / 4) MCP_DBG_CHECK(reinterpret_cast<unsigned int*>(after_ring - MCP_DBG_CANARY_BYTES)[tid] == MCP_DBG_CANARY, DBG_RING_CANARY);
#endif
    // never leave with bulk copies in flight into this CTA's shared memory
    if (tid == 0) {
        while (cv.n < cs.n) issue_carry();  // (stop path) complete the armed barriers
        for (; g < cs.n; ++g) {
            while (!mbar_try_wait(full + cst, cpar)) {}
            if (++cst == n_stages) { cst = 0; cpar ^= 1u; }
        }
    }
    __syncthreads();
}

template <int P, bool TAU>
__global__ void __launch_bounds__(NT, 1) lsm_persist_kernel(Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (blockIdx.x == 0) {
        // reducer: the dynamic shared memory is its gather area
        unsigned int* got = reinterpret_cast<unsigned int*>(smem_raw);
        double* vals = reinterpret_cast<double*>(smem_raw + RED_GOT_BYTES);
        reducer<P>(a, got, vals);
    } else {
        worker<P, TAU>(a, smem_raw);
    }
}

}  // namespace px
