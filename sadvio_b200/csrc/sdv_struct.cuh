// Structure pass of an upload ON THE DEVICE (round 2): the slot lists — (landmark, distinct keyframe) groups in first-appearance
// order, the grouping addResidualsLocalMap's walk induces (AngularAdjustmentCERESAnalytic.cpp:247-289) — and the tiles of the
// fused kernels are derived from the observation arrays the bulk arena carries anyway, instead of being built by four host
// threads, packed and copied (0.15-0.2 ms of a 0.9 ms end-to-end solve at C3, and 0.7 MB of H2D).  The host keeps what it needs
// to choose kernels before anything runs: range validation, the keyframes in use, slot count, the per-landmark keyframe span
// that bounds the band of the reduced system.
//
// Precondition (checked here, `flags[0]`): the observations of a landmark are grouped by keyframe (stereo pairs adjacent — what
// the reference's feature lists look like); a keyframe that re-appears after another one makes the host fall back to its own
// general grouping.
#pragma once
#include "sdv_types.cuh"

namespace sdv {

constexpr int SCAN_T = 256, SCAN_PER = 4, SCAN_BLOCK = SCAN_T * SCAN_PER; // elements per CTA of the scans

// head[o] = 1 when observation o opens a slot (first observation of its landmark, or another keyframe than the one before)
// Every kernel reads its sizes and buffers from the device-resident problem description, and the grids are sized from the
// handle's capacities: the eight launches are captured ONCE per handle into a small CUDA graph and replayed per upload.
__global__ void k_struct_heads(const DevProblem *__restrict__ Pg) {
    const DevProblem &P = *Pg;
    if (!P.st_on) return;
    const int *__restrict__ obs_lmk = P.obs_lmk, *__restrict__ obs_fc = P.obs_fc;
    const int C = P.C, o0 = P.o0, Oloc = P.o1 - P.o0;
    int *head = P.st_head, *flags = P.st_tot + 8;
    for (int ol = blockIdx.x * blockDim.x + threadIdx.x; ol < Oloc; ol += gridDim.x * blockDim.x) {
        const int o = o0 + ol;
        const int l = obs_lmk[o], f = obs_fc[o] / C;
        const bool head_l = ol == 0 || obs_lmk[o - 1] != l;
        const bool head_s = head_l || obs_fc[o - 1] / C != f;
        head[ol] = head_s ? 1 : 0;
        if (head_s && !head_l) { // the keyframe must not have appeared earlier in this landmark's list
            for (int q = o - 2; q >= o0 && obs_lmk[q] == l; q--)
                if (obs_fc[q] / C == f) {
                    flags[0] = 1;
                    break;
                }
        }
    }
}

// exclusive scan, three launches: per-CTA scan + CTA totals, scan of the totals (one CTA), and the consumers add the CTA offset
// which = 0: the slot heads (n = local observations); 1: the tile heads (n = landmarks of this rank)
__global__ void __launch_bounds__(SCAN_T) k_scan_blocks(const DevProblem *__restrict__ Pg, int which) {
    const DevProblem &P = *Pg;
    if (!P.st_on) return;
    const int *__restrict__ in = which ? P.st_thead : P.st_head;
    int *out = which ? P.st_tidx : P.st_sidx, *sums = which ? P.st_tsum : P.st_ssum;
    const int n = which ? P.l1 - P.l0 : P.o1 - P.o0;
    if (blockIdx.x * SCAN_BLOCK >= n) return; // (grids are sized from capacities)
    __shared__ int wsum[SCAN_T / 32];
    const int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_PER;
    int v[SCAN_PER], tot = 0;
#pragma unroll
    for (int q = 0; q < SCAN_PER; q++) {
        v[q] = base + q < n ? in[base + q] : 0;
        tot += v[q];
    }
    int incl = tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_T / 32 ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += x;
        }
        if (lane < SCAN_T / 32) wsum[lane] = s;
    }
    __syncthreads();
    int run = incl - tot + (w > 0 ? wsum[w - 1] : 0);
#pragma unroll
    for (int q = 0; q < SCAN_PER; q++) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
    if (threadIdx.x == SCAN_T - 1) sums[blockIdx.x] = wsum[SCAN_T / 32 - 1];
}
__global__ void __launch_bounds__(1024) k_scan_sums(const DevProblem *__restrict__ Pg, int which) {
    const DevProblem &P = *Pg;
    if (!P.st_on) return;
    int *sums = which ? P.st_tsum : P.st_ssum, *total = P.st_tot + which;
    const int n = which ? P.l1 - P.l0 : P.o1 - P.o0, nblk = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nblk; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nblk ? sums[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            int s = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += x;
            }
            wsum[lane] = s;
        }
        __syncthreads();
        const int excl = carry + incl - v + (w > 0 ? wsum[w - 1] : 0);
        if (i < nblk) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry += wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// slot arrays from the scanned heads; slot_obs is the identity (observations are already grouped by slot)
__global__ void k_struct_slots(const DevProblem *__restrict__ Pg) {
    const DevProblem &P = *Pg;
    if (!P.st_on) return;
    const int *__restrict__ obs_lmk = P.obs_lmk, *__restrict__ obs_fc = P.obs_fc, *__restrict__ head = P.st_head, *__restrict__ sidx = P.st_sidx, *__restrict__ ssum = P.st_ssum;
    const int *__restrict__ total = P.st_tot;
    const int C = P.C, o0 = P.o0, Oloc = P.o1 - P.o0, L = P.L, l0 = P.l0, l1 = P.l1;
    int *slot_ptr = const_cast<int *>(P.slot_ptr), *slot_frame = const_cast<int *>(P.slot_frame), *slot_obs_ptr = const_cast<int *>(P.slot_obs_ptr),
        *slot_obs = const_cast<int *>(P.slot_obs);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const int nslots = *total;
    for (int ol = tid; ol < Oloc; ol += nt) {
        slot_obs[ol] = ol;
        if (head[ol]) {
            const int s = sidx[ol] + ssum[ol / SCAN_BLOCK];
            slot_frame[s] = obs_fc[o0 + ol] / C;
            slot_obs_ptr[s] = ol;
        }
    }
    if (tid == 0) slot_obs_ptr[nslots] = Oloc;
    // first slot of every landmark: the slot of its first observation (landmarks without observations take the next one's)
    for (int l = tid; l <= L; l += nt) {
        int sp;
        if (l <= l0) sp = 0;
        else if (l >= l1) sp = nslots;
        else {
            int lo = 0, hi = Oloc; // first local observation with obs_lmk >= l
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (obs_lmk[o0 + mid] < l) lo = mid + 1;
                else hi = mid;
            }
            sp = lo < Oloc ? sidx[lo] + ssum[lo / SCAN_BLOCK] : nslots; // (lo is the first observation of a landmark: a head)
        }
        slot_ptr[l] = sp;
    }
}

// tiles: landmarks whose first slot falls into the same bucket of `capq` slots form a tile (<= capq - 1 + max_slots <= cap slots),
// cut again every FT_LMK landmarks.  thead[l - l0] = 1 when landmark l opens a tile.
__global__ void k_struct_tile_heads(const DevProblem *__restrict__ Pg, int max_lmk) {
    const DevProblem &P = *Pg;
    if (!P.st_on) return;
    const int *__restrict__ slot_ptr = P.slot_ptr;
    const int l0 = P.l0, l1 = P.l1, capq = P.st_capq;
    int *thead = P.st_thead;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < l1 - l0; i += gridDim.x * blockDim.x) {
        const int l = l0 + i;
        const int b = slot_ptr[l] / capq;
        bool start = i == 0 || slot_ptr[l - 1] / capq != b;
        if (!start) {
            int lo = l0, hi = l; // first landmark of this bucket
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (slot_ptr[mid] / capq < b) lo = mid + 1;
                else hi = mid;
            }
            start = (l - lo) % max_lmk == 0;
        }
        thead[i] = start ? 1 : 0;
    }
}
__global__ void k_struct_tiles(DevProblem *Pw) {
    const DevProblem &Pr = *Pw;
    if (!Pr.st_on) return;
    const int *__restrict__ thead = Pr.st_thead, *__restrict__ tidx = Pr.st_tidx, *__restrict__ tsum = Pr.st_tsum, *__restrict__ total = Pr.st_tot + 1, *__restrict__ nslots = Pr.st_tot;
    const int l0 = Pr.l0, l1 = Pr.l1;
    int *tile_ptr = const_cast<int *>(Pr.tile_ptr);
    DevProblem *P = Pw;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = tid; i < l1 - l0; i += nt)
        if (thead[i]) tile_ptr[tidx[i] + tsum[i / SCAN_BLOCK]] = l0 + i;
    if (tid == 0) {
        const int ntiles = l1 > l0 ? *total : 0;
        tile_ptr[ntiles] = l1;
        P->ntiles = ntiles;
        P->nslots = *nslots;
    }
}

} // namespace sdv
