// lsm_persist.cuh -- the whole Longstaff-Schwartz backward induction (LSMPricer.cpp:35-101) in ONE cooperative launch.
// Included by lsm.cu after the packed-fp32 step arithmetic (fast2_compute) and the bulk-copy helpers.
//
// Why.  One launch per time step costs, per step, a kernel boundary plus a serial tail: per-CTA partials -> ticket ->
// last CTA folds -> (multi-GPU) exchange -> one thread solves -> only then may the dependent launch proceed.  At 2^23
// paths per GPU (config 3 on 8 GPUs) that tail is a third of the step.  Here the grid stays resident for all M steps:
//
//   worker CTAs (blockIdx 1 .. nw)   own a FIXED set of 4096-path tiles for the whole induction and stream them through a
//       TMA ring that never drains: the tile sequence (step s, tile it) is one continuous stream, and because a
//       worker's carry tiles are written by that worker only, the loads of step s+1 are issued while step s is still
//       being computed (generic-proxy stores -> fence.proxy.async -> mbarrier -> bulk copy).  Per step a worker sends
//       ONE row of regression moments to the reducer and waits for ONE row of coefficients.
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
    int is_call, M, ns, l2_resident, n_workers, n_stages;
    McpPx x;
    unsigned long long seq0;  // tag of this launch's first exchange; exchange e uses seq0 + e
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
__device__ __forceinline__ unsigned long long* local_bc(const McpPx& x, unsigned long long seq) {
    return x.local + (size_t)2 * MCP_PX_MAXW * MCP_PX_ROWW + (size_t)(seq & 1ull) * MCP_PX_BCW;
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
        do {
            w = ld_word(p);
            if ((w & 0xffffffff00000000ull) == tag) break;
            if (global_ns() - t0 > PX_TIMEOUT_NS) return false;
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
            while (true) {
                all = true;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if ((w[u] & 0xffffffff00000000ull) != tag) w[u] = ld_word(p[u]);
                    all = all && ((w[u] & 0xffffffff00000000ull) == tag);
                }
                if (all) break;
                if (global_ns() - t0 > PX_TIMEOUT_NS || *(volatile int*)s_fail) { *(volatile int*)s_fail = 1; break; }
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

// status word of a broadcast: 0 = go on, 1 = stop (error)
__device__ __forceinline__ void broadcast(const McpPx& x, unsigned long long seq, const double* vals, int nv, int status) {
    const unsigned long long tag = tag_of(seq);
    unsigned long long* bc = local_bc(x, seq);
    for (int k = threadIdx.x; k < 2 * nv; k += NT) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(vals[k >> 1]);
        st_word(bc + 1 + k, ((k & 1) ? (b >> 32) : (b & 0xffffffffull)) | tag);
    }
    if (threadIdx.x == 0) st_word(bc, (unsigned long long)(unsigned int)status | tag);
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
        broadcast(a.x, at_seq, vals, 0, 1);
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
        broadcast(a.x, seq_bc, vals, 0, 0);
    }

    // ---- the induction: one exchange per step that regresses, one for {sum V0, N}, one for the squared deviations ----
    for (int j = M - 1; j >= 0; --j) {
        const bool dm = j > 0 && a.kind[j - 1] == 0, df = j == 0;
        if (!dm && !df) continue;
        const unsigned long long tag = tag_of(seq);
        ok = gather_rows(local_row(a.x, seq, 0), MCP_PX_ROWW, nw, 2 * NV, tag, got, &s_fail);
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
        ok = exchange(a.x, seq, vals, NV, got, &s_fail, ok ? 0 : 1) && ok;
        if (!ok) { give_up(seq); return; }
        if (df) {
            if (tid == 0) {
                a.fin[0] = vals[0];
                a.fin[2] = vals[1];
                s_coef[0] = vals[0] / vals[1];  // the mean: workers need it for the squared deviations
            }
            __syncthreads();
            broadcast(a.x, seq, s_coef, 1, 0);
        } else {
            if (tid == 0) solve_normal_equations<P>(vals, s_coef);
            __syncthreads();
            if (tid < COEF_LD) a.coef[(size_t)(j - 1) * COEF_LD + tid] = s_coef[tid];
            broadcast(a.x, seq, s_coef, P + 1, 0);
        }
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
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)n_stages * STAGE_BYTES);
    uint64_t* empty = full + 8;
    double* sacc = reinterpret_cast<double*>(smem_raw + (size_t)n_stages * STAGE_BYTES + 128);  // [NV][NT]
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
    const uint64_t pol_first = l2_policy_evict_first(), pol_norm = l2_policy_evict_normal();
    unsigned long long seq = a.seq0;

    // tile g of the stream = (step s = g / my_tiles, tile it = g % my_tiles); one elected thread arms the stage's barrier
    // with the bytes of all its copies and starts them
    auto issue = [&](int64_t g) {
        const int s = (int)(g / my_tiles);
        const int64_t it = g - (int64_t)s * my_tiles;
        const int j = M - 1 - s;
        const bool dm = j > 0 && __ldg(a.kind + j - 1) == 0, want_v = s > 0;
        const int st = (int)(g % n_stages);
        const int64_t i0 = ((int64_t)w + it * nw) * TILE;
        const int64_t cnt = a.ld - i0 < TILE ? a.ld - i0 : TILE;  // rows are padded to ld (multiple of 128)
        const uint32_t bytes = (uint32_t)cnt * 4u;
        float* dst = ring + (size_t)st * STAGE_FLOATS;
        mbar_expect_tx(full + st, bytes * (1u + (dm ? 1u : 0u) + (want_v ? 1u : 0u)));
        bulk_g2s_hint(dst, S + (int64_t)j * a.ld + i0, bytes, full + st, pol_first);              // last use of row j
        if (dm) bulk_g2s_hint(dst + TILE, S + (int64_t)(j - 1) * a.ld + i0, bytes, full + st, pol_norm);  // read again next step
        if (want_v) bulk_g2s_hint(dst + 2 * TILE, V + i0, bytes, full + st, a.l2_resident ? pol_norm : pol_first);
    };
    int64_t issued = 0;  // thread 0 only
    if (tid == 0) {
        s_stop = 0;
        for (int st = 0; st < n_stages; ++st) { mbar_init(full + st, 1); mbar_init(empty + st, NT / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the terminal step reads no carry: its tiles can fly before anything else happens
        for (; issued < n_stages && issued < my_tiles; ++issued) issue(issued);
    }
    for (int m = 0; m < NV; ++m) sacc[m * NT + tid] = 0.0;

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
        if (!wait_word(local_bc(a.x, seq_tables), tag_of(seq_tables), &status) || status != 0) s_stop = 1;
    }
    __syncthreads();
    __threadfence();

    FastConsts<P> k;
    SweepArgs sa;  // the view fast2_compute expects
    memset(&sa, 0, sizeof(sa));
    sa.n = a.n; sa.tau = a.tau; sa.is_call = a.is_call; sa.K = a.K; sa.disc = a.disc;
    {
        const float sgn = a.is_call ? 1.f : -1.f;
        const float K_hi = (float)a.K, K_lo = (float)(a.K - (double)K_hi);
        k.sg = splat2(sgn);
        k.nsK = splat2(-sgn * K_hi);
        k.nsKlo = splat2(-sgn * K_lo);
        const float d_hi = (float)a.disc;
        k.d_hi = splat2(d_hi);
        k.d_lo = splat2((float)(a.disc - (double)d_hi));
    }
    float2 la[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) la[m] = make_float2(0.f, 0.f);
    int since = 0;
    int64_t g = 0;
    unsigned long long seq_coef = 0;  // sequence number whose broadcast carries the coefficients of the coming step
    double mean = 0.0;

    for (int s = 0; s < M && !s_stop; ++s) {
        const int j = M - 1 - s;
        const int mode = s == 0 ? 2 : __ldg(a.kind + j);
        const bool dm = j > 0 && __ldg(a.kind + j - 1) == 0, df = j == 0;
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();  // everybody is done with the previous step's constants (and, short streams, with its carry stores)
        if (!chain && s > 0 && tid == 0)
            for (; issued < (int64_t)(s + 1) * my_tiles; ++issued) issue(issued);
        // ---- constants of the step: c_j from the reducer's broadcast (the only wait of the step), mu / 1/s from the tables ----
        if (tid < 4) {
            const int jj = (tid < 2) ? j : (j > 0 ? j - 1 : 0);
            s_c[COEF_LD + tid] = __ldcg(((tid & 1) ? a.inv_s : a.mu) + jj);
        }
        if (mode == 0) {
            if (tid <= 2 * (P + 1)) {  // word 0 = status, words 1.. = c_0 .. c_P as (low, high) halves
                unsigned int lo = 1;
                const bool got_it = wait_word(local_bc(a.x, seq_coef) + tid, tag_of(seq_coef), &lo);
                if (tid == 0) { if (!got_it || lo != 0) s_stop = 1; }
                else reinterpret_cast<unsigned int*>(s_c)[tid - 1] = got_it ? lo : 0u;
            }
        }
        __syncthreads();
        if (s_stop) break;
#pragma unroll
        for (int m = 0; m <= P; ++m) k.c[m] = splat2(mode == 0 ? (float)s_c[m] : 0.f);
        k.nmu = splat2(-(float)s_c[COEF_LD]);
        k.is = splat2((float)s_c[COEF_LD + 1]);
        k.nmu_p = splat2(-(float)s_c[COEF_LD + 2]);
        k.is_p = splat2((float)s_c[COEF_LD + 3]);
        sa.j = j; sa.terminal = s == 0; sa.do_moments = dm; sa.do_final = df;

        auto run_tiles = [&](auto kind_tag) {
            constexpr int KIND = decltype(kind_tag)::value;
            const bool ldm = KIND == 0 ? true : dm, wv = KIND == 0 ? true : (mode != 2);
            for (int64_t it = 0; it < my_tiles; ++it, ++g) {
                const int st = (int)(g % n_stages);
                const uint32_t parity = (uint32_t)((g / n_stages) & 1);
                while (!mbar_try_wait(full + st, parity)) {}
                const float* buf = ring + (size_t)st * STAGE_FLOATS;
                F8 s8, p8, v8;
                {
                    const float4 x0 = *reinterpret_cast<const float4*>(buf + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + 2048 + 4 * tid);
                    s8.q[0] = make_float2(x0.x, x0.y); s8.q[1] = make_float2(x0.z, x0.w); s8.q[2] = make_float2(x1.x, x1.y); s8.q[3] = make_float2(x1.z, x1.w);
                }
                p8 = s8; v8 = s8;
                if (ldm) {
                    const float4 x0 = *reinterpret_cast<const float4*>(buf + TILE + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + TILE + 2048 + 4 * tid);
                    p8.q[0] = make_float2(x0.x, x0.y); p8.q[1] = make_float2(x0.z, x0.w); p8.q[2] = make_float2(x1.x, x1.y); p8.q[3] = make_float2(x1.z, x1.w);
                }
                if (wv) {
                    const float4 x0 = *reinterpret_cast<const float4*>(buf + 2 * TILE + 4 * tid), x1 = *reinterpret_cast<const float4*>(buf + 2 * TILE + 2048 + 4 * tid);
                    v8.q[0] = make_float2(x0.x, x0.y); v8.q[1] = make_float2(x0.z, x0.w); v8.q[2] = make_float2(x1.x, x1.y); v8.q[3] = make_float2(x1.z, x1.w);
                }
                // Hand the stage back.  Everything this thread stored to the carry so far (tiles < g of the stream) is
                // ordered before the arrival and made visible to the async proxy, so when all 16 warps have arrived the
                // elected thread may start the copies of stream tile g + n_stages -- whose carry tile was written at
                // stream position g + n_stages - my_tiles < g (the host guarantees my_tiles > n_stages).
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + st);
                if (tid == 0 && chain && g + n_stages < g_total) {
                    while (!mbar_try_wait(empty + st, parity)) {}
                    issue(g + n_stages);
                    issued = g + n_stages + 1;
                }
                const int64_t i0 = ((int64_t)w + it * nw) * TILE, ia = i0 + 4 * tid, ib = i0 + 2048 + 4 * tid;
                if (i0 + TILE <= a.n) fast2_compute<P, TAU, false, KIND>(sa, k, s8, p8, v8, ia, ib, mode, la);
                else fast2_compute<P, TAU, true, KIND>(sa, k, s8, p8, v8, ia, ib, mode, la);
                if (a.l2_resident) {
                    if (ia < a.ld) stg4_keep(V + ia, v8.q[0], v8.q[1]);
                    if (ib < a.ld) stg4_keep(V + ib, v8.q[2], v8.q[3]);
                } else {
                    if (ia < a.ld) stg4_stream(V + ia, v8.q[0], v8.q[1]);
                    if (ib < a.ld) stg4_stream(V + ib, v8.q[2], v8.q[3]);
                }
                if (++since == FLUSH) {
#pragma unroll
                    for (int m = 0; m < NV; ++m) {
                        sacc[m * NT + tid] += (double)(la[m].x + la[m].y);
                        la[m] = make_float2(0.f, 0.f);
                    }
                    since = 0;
                }
            }
        };
        if (mode == 0 && dm && !df) run_tiles(std::integral_constant<int, 0>{});
        else run_tiles(std::integral_constant<int, 1>{});

        if (dm || df) {
            // ---- this worker's row of the step: warp sums -> 16 rows in shared memory -> thread m adds them in warp order ----
#pragma unroll
            for (int m = 0; m < NV; ++m) {
                const double t = warp_sum(sacc[m * NT + tid] + (double)(la[m].x + la[m].y));
                if (lane == 0) red[warp][m] = t;
                sacc[m * NT + tid] = 0.0;
                la[m] = make_float2(0.f, 0.f);
            }
            since = 0;
            __syncthreads();
            if (tid < NV) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < NT / 32; ++q) t += red[q][tid];
                st_tagged_double(local_row(a.x, seq, w), tid, t, tag_of(seq));
            }
            seq_coef = seq;
            ++seq;
        }
    }

    // ---- standard error: the reducer's last broadcast carried the global mean ----
    if (!s_stop) {
        if (tid < 3) {
            unsigned int lo = 1;
            const bool got_it = wait_word(local_bc(a.x, seq_coef) + tid, tag_of(seq_coef), &lo);
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
    // never leave with bulk copies in flight into this CTA's shared memory
    if (tid == 0)
        for (; g < issued; ++g) {
            const int st = (int)(g % n_stages);
            while (!mbar_try_wait(full + st, (uint32_t)((g / n_stages) & 1))) {}
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
