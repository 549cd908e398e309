// Multi-GPU exchange over NVLink / NVSwitch peer memory: the per-iteration all-reduce of the band of the
// reduced system, the 7 candidate scalars and the landmark gather of the epilogue as ONE kernel each, instead of
// pack -> ncclAllReduce -> unpack.  At F = 50 the payload is 0.5 MB: the exchange is launch- and latency-bound, and an NCCL
// call costs more in launch + protocol than the wire time; a plain kernel is also just another node of the whole-solve CUDA
// graph (conditional WHILE body included), which NCCL calls are not.
//
// Every rank owns an exchange area (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the peers):
//     data  [2 parities][world sources][cap] doubles      written by the PEERS (st.global over NVLink), read locally
//     flags [2 parities][world sources][PEER_MAX_CHUNKS]  epoch numbers, written by the peers with st.release.sys
//     epoch [0] exchanges completed on this rank, [1] CTA ticket, [2] timeout flag
// One-shot all-reduce, chunked by CTA: CTA c pushes its chunk of the local values into every peer's data[parity][me], publishes
// flag[parity][me][c] = epoch, waits for the same flag from every peer, and sums the world copies IN RANK ORDER — every rank
// adds the same numbers in the same order, so the replicated LM state stays bit-identical across ranks.  No CTA waits for
// another CTA of its own grid.  Two parities suffice: a rank finishes exchange e+1 only after every peer has started it,
// i.e. finished reading exchange e, so nobody can be overwritten two exchanges ahead.
#pragma once
#include "sdv_kernels.cuh"

namespace sdv {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_MAX_CHUNKS = 256;
constexpr int PEER_THREADS = 256;

struct PeerXchg { // kernel argument: fixed once the peers' areas are open, so captured launches stay valid
    int rank, world;
    unsigned long long cap;                      // doubles per (parity, source) slot
    double *data[PEER_MAX_WORLD];                // data area of every rank (own one included)
    unsigned long long *flags[PEER_MAX_WORLD];   // flag area of every rank
    unsigned long long *epoch;                   // this rank's [epoch, ticket, timeout]
};

SDV_DEV void st_release_sys(unsigned long long *p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
SDV_DEV unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SDV_DEV unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// what is exchanged: MODE 0 a contiguous buffer; MODE 1 the band of S + [g | diag | raw gradient] + the gradient-tolerance flag of
// this rank's landmark columns (the layout k_band_exchange packs for the NCCL path); MODE 2 the candidate scalars of Accum
struct PeerMap {
    double *buf;       // MODE 0: the buffer; MODE 1: Sb
    int n_pad, ld, W;  // MODE 1
    Accum *acc;        // MODE 1 / 2
    const LMState *st; // MODE 2
    double grad_tol;   // MODE 1
    int which;         // MODE 2: linearisation buffer whose cost is exchanged (-2: the candidate)
};

template <int MODE> SDV_DEV double peer_load(const PeerMap &m, size_t e, bool &real) {
    real = true;
    if (MODE == 0) return m.buf[e];
    if (MODE == 1) {
        const size_t nband = (size_t)m.n_pad * m.W, ntot = nband + 3 * (size_t)m.n_pad;
        if (e < nband) {
            const int i = (int)(e / m.W), c = (int)(e - (size_t)i * m.W), col = i - m.W + 1 + c;
            if (col < 0) {
                real = false;
                return 0.0;
            }
            return m.buf[(size_t)i * m.ld + col];
        }
        if (e < ntot) return m.buf[(size_t)m.n_pad * m.ld + (e - nband)];
        return (__longlong_as_double((long long)m.acc->grad_max_bits) > m.grad_tol) ? 1.0 : 0.0;
    }
    const int b = m.which >= 0 ? m.which : 1 - m.st->cur;
    switch ((int)e) {
    case 0: return m.acc->cost[b];
    case 1: return m.acc->model_gd;
    case 2: return m.acc->model_dd;
    case 3: return m.acc->step_norm2;
    case 4: return m.acc->cand_norm2;
    case 5: return m.which == 0 ? m.acc->fixed_cost : 0.0;
    default: return m.acc->schur_fail ? 1.0 : 0.0;
    }
}
template <int MODE> SDV_DEV void peer_store(const PeerMap &m, size_t e, double v) {
    if (MODE == 0) {
        m.buf[e] = v;
        return;
    }
    if (MODE == 1) {
        const size_t nband = (size_t)m.n_pad * m.W, ntot = nband + 3 * (size_t)m.n_pad;
        if (e < nband) {
            const int i = (int)(e / m.W), c = (int)(e - (size_t)i * m.W), col = i - m.W + 1 + c;
            m.buf[(size_t)i * m.ld + col] = v;
        } else if (e < ntot) {
            m.buf[(size_t)m.n_pad * m.ld + (e - nband)] = v;
        } else {
            m.acc->grad_max_bits = v > 0.0 ? (unsigned long long)__double_as_longlong(1e300) : 0ull;
        }
        return;
    }
    const int b = m.which >= 0 ? m.which : 1 - m.st->cur;
    switch ((int)e) {
    case 0: m.acc->cost[b] = v; break;
    case 1: m.acc->model_gd = v; break;
    case 2: m.acc->model_dd = v; break;
    case 3: m.acc->step_norm2 = v; break;
    case 4: m.acc->cand_norm2 = v; break;
    case 5: if (m.which == 0) m.acc->fixed_cost = v; break;
    default: m.acc->schur_fail = v > 0.0 ? 1 : 0; break;
    }
}

template <int MODE> __global__ void __launch_bounds__(PEER_THREADS) k_peer_allreduce(PeerXchg X, PeerMap m, unsigned long long count) {
    __shared__ unsigned long long s_epoch;
    const int t = threadIdx.x, me = X.rank, world = X.world;
    if (t == 0) s_epoch = *reinterpret_cast<volatile unsigned long long *>(X.epoch);
    __syncthreads();
    const unsigned long long ep = s_epoch + 1;
    const unsigned long long par = ep & 1ull;
    const unsigned long long per = ((count + gridDim.x - 1) / gridDim.x + 31ull) & ~31ull;
    const unsigned long long e0 = per * blockIdx.x, e1 = e0 + per < count ? e0 + per : count;
    if (e0 < e1) {
        // ---- push this chunk into every peer's slot [parity][me]
        for (unsigned long long e = e0 + t; e < e1; e += PEER_THREADS) {
            bool real;
            const double v = peer_load<MODE>(m, e, real);
            for (int p = 0; p < world; p++)
                if (p != me) X.data[p][(par * world + me) * X.cap + e] = v;
        }
        __syncthreads();
        if (t < world && t != me) {
            __threadfence_system();
            st_release_sys(X.flags[t] + (par * world + me) * PEER_MAX_CHUNKS + blockIdx.x, ep);
        }
        // ---- wait for the same chunk of every peer (bounded: a missing peer must not hang the GPU)
        if (t < world && t != me) {
            const unsigned long long *f = X.flags[me] + (par * world + t) * PEER_MAX_CHUNKS + blockIdx.x;
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(f) < ep) {
                if (global_timer_ns() - t0 > 30000000000ull) { // 30 s: ranks may reach a solve seconds apart; a dead peer must not hang the GPU for ever
                    X.epoch[2] = 1ull;
                    break;
                }
            }
        }
        __syncthreads();
        // ---- sum the world copies in rank order (identical on every rank)
        const double *mine = X.data[me] + par * world * X.cap;
        for (unsigned long long e = e0 + t; e < e1; e += PEER_THREADS) {
            bool real;
            const double own = peer_load<MODE>(m, e, real);
            if (!real) continue;
            double v = 0.0;
            for (int s = 0; s < world; s++) v += s == me ? own : __ldcg(mine + s * X.cap + e);
            peer_store<MODE>(m, e, v);
        }
    }
    // ---- the last CTA closes the exchange
    __syncthreads();
    if (t == 0) {
        __threadfence();
        const unsigned long long ticket = atomicAdd(X.epoch + 1, 1ull);
        if (ticket == gridDim.x - 1) {
            X.epoch[1] = 0ull;
            __threadfence();
            *reinterpret_cast<volatile unsigned long long *>(X.epoch) = ep;
        }
    }
}

} // namespace sdv
