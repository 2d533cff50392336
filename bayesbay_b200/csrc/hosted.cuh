// Mapped hosted step: the chains live in PINNED HOST memory as fixed-stride records and the chain kernel itself reads
// each chain's uploaded state and writes its new state over PCIe (zero-copy), so that the host <-> device exchange of
// the reference's `Sampler.advance_chain` round trip (samplers/_samplers.py:225-237: every chain pickled to the pool,
// advanced, pickled back) overlaps the step instead of bracketing it.
//
// Record of chain c: `stride` 8-byte words at  image + c * stride:
//   header (hrw words): k of every space (int64) | sigma (and correlation) per hierarchical target | temperature |
//                       iteration (int64);
//   space s: cells at word base[s] + j * (ndim + n_params) + q,  j < k_s  (site coordinates, then parameter values;
//            the record has room for kmax cells, only the k in use are read / written -- only they cross the bus).
#pragma once
#include "common.cuh"

namespace bbb {

struct HostedIO {
    const unsigned long long *in;  // uploaded records, mapped pinned host memory; nullptr: not a hosted step
    unsigned long long *out;       // records of the new states, mapped pinned host memory
    int32_t *mismatch;             // device [C]: set for a chain whose uploaded state is not bit-identical to the resident one
    int32_t *host_flag;            // mapped pinned word: set when some chain mismatched
    const int32_t *only;           // device [C] or nullptr: only chains with a non-zero word run (the redo pass)
    unsigned long long *moved;     // device [2]: bytes read from / written to host memory (statistics)
    // developer switch BBB_MAPPED_PDL=1 (measured slower, see abi.cu): chain_done[c] = seq once chain c's new STATE is
    // complete -- what the next proposal's warp for chain c then waits for, the proposal kernel being launched as this
    // kernel's programmatic dependent
    int32_t *chain_done;           // device [C] or nullptr
    int seq;
    long long stride;              // words per record
    int hrw;                       // header words
    int base[BBB_MAX_SPACES];      // first cell word of each space
};

__device__ __forceinline__ unsigned long long hosted_bits(double v) { return (unsigned long long)__double_as_longlong(v); }

// header word i of chain c's CURRENT state
__device__ __noinline__ inline unsigned long long hosted_header_word(const bbb_problem_spec &P, const bbb_buffers &B, long long c, int i) {
    if (i < P.n_spaces) return (unsigned long long)(long long)k_ref(B, i, (int)B.sel[i][c], c);
    int rel = i - P.n_spaces;
    for (int t = 0; t < P.n_targets; ++t) {
        const bbb_target_spec &T = P.targets[t];
        if (T.cov_kind != BBB_COV_HIERARCHICAL) continue;
        const int n = T.correlated ? 2 : 1;
        if (rel < n) return hosted_bits(rel == 0 ? B.sigma[t][c] : B.corr[t][c]);
        rel -= n;
    }
    return rel == 0 ? hosted_bits(B.temperature[c]) : (unsigned long long)B.iteration[c];
}
// cell word i (= j * w + q) of space s, slot `slot`
__device__ __forceinline__ unsigned long long hosted_cell_word(const bbb_buffers &B, const bbb_space_spec &S, int s, int slot, long long c, int i) {
    const int w = S.ndim + S.n_params, j = i / w, q = i - j * w;
    return hosted_bits(q < S.ndim ? site_ref(B, S, s, slot, q, j, c) : value_ref(B, S, s, slot, q - S.ndim, j, c));
}

// This thread's share (words tid, tid + nth, ...) of "uploaded record == resident current state"; true on a mismatch.
// `rec`: the chain's uploaded record -- in host memory (io.in + c * io.stride) or its staged copy in shared memory.
__device__ __forceinline__ bool hosted_compare(const EngineParams &ep, const HostedIO &io, const unsigned long long *rec, long long c, int tid,
                                               int nth) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    bool bad = false;
    for (int i = tid; i < io.hrw; i += nth) bad |= rec[i] != hosted_header_word(P, B, c, i);
    for (int s = 0; s < P.n_spaces; ++s) {
        const bbb_space_spec &S = P.spaces[s];
        const int slot = (int)B.sel[s][c], n = k_ref(B, s, slot, c) * (S.ndim + S.n_params);
        for (int i = tid; i < n; i += nth) bad |= rec[io.base[s] + i] != hosted_cell_word(B, S, s, slot, c, i);
    }
    return bad;
}
// The same with this thread's FIRST word of the record (word `tid` of the contiguous run header + cells of space 0) already
// in a register: the chain kernel requests it when the block starts, so that the trip over the bus overlaps the tables
// and the first rasterisation pass.  `pref` is rec[tid] when tid < hrw + (cells of space 0 in use) words.
__device__ __forceinline__ bool hosted_compare_prefetched(const EngineParams &ep, const HostedIO &io, const unsigned long long *rec, long long c,
                                                          int tid, int nth, unsigned long long pref) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    bool bad = false;
    {
        const bbb_space_spec &S = P.spaces[0];
        const int slot = (int)B.sel[0][c], n = io.hrw + k_ref(B, 0, slot, c) * (S.ndim + S.n_params);
        for (int j = tid; j < n; j += nth) {
            const unsigned long long up = j == tid ? pref : rec[j];
            const unsigned long long here = j < io.hrw ? hosted_header_word(P, B, c, j) : hosted_cell_word(B, S, 0, slot, c, j - io.hrw);
            bad |= up != here;
        }
    }
    for (int s = 1; s < P.n_spaces; ++s) {
        const bbb_space_spec &S = P.spaces[s];
        const int slot = (int)B.sel[s][c], n = k_ref(B, s, slot, c) * (S.ndim + S.n_params);
        for (int i = tid; i < n; i += nth) bad |= rec[io.base[s] + i] != hosted_cell_word(B, S, s, slot, c, i);
    }
    return bad;
}
// The CURRENT state of chain c into its record of the out image (all threads of the group; the caller has made the
// decision's writes visible with a barrier).  Thread 0 also counts the bytes of the exchange.
// (s_known >= 0: the caller knows the current slot and k of that space -- the chain kernel after its decision -- which
// saves the two dependent loads in front of the cells' loads)
__device__ __forceinline__ void hosted_write(const EngineParams &ep, const HostedIO &io, long long c, int tid, int nth, int words_in,
                                             int s_known = -1, int slot_known = 0, int k_known = 0) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    unsigned long long *rec = io.out + c * io.stride;
    for (int i = tid; i < io.hrw; i += nth)
        rec[i] = i == s_known ? (unsigned long long)(long long)k_known : hosted_header_word(P, B, c, i);
    int words = io.hrw;
    for (int s = 0; s < P.n_spaces; ++s) {
        const bbb_space_spec &S = P.spaces[s];
        const int slot = s == s_known ? slot_known : (int)B.sel[s][c];
        const int n = (s == s_known ? k_known : k_ref(B, s, slot, c)) * (S.ndim + S.n_params);
        for (int i = tid; i < n; i += nth) rec[io.base[s] + i] = hosted_cell_word(B, S, s, slot, c, i);
        words += n;
    }
    if (tid == 0 && io.moved != nullptr) {
        atomicAdd(io.moved, 8ull * (unsigned long long)words_in);
        atomicAdd(io.moved + 1, 8ull * (unsigned long long)words);
    }
}
// Completion signal of chain c's step (thread 0 of its block, right after the decision's last write of the chain's
// STATE): released at device scope for the next iteration's proposal warp of the chain (k_propose, step_kernels.cu).
__device__ __forceinline__ void hosted_signal_done(const HostedIO &io, long long c) {
    if (io.chain_done == nullptr) return;
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(io.chain_done + c), "r"(io.seq) : "memory");
}
// words of chain c's record that are in use (header + the cells of every space), from the resident state
__device__ __forceinline__ int hosted_words_in_use(const EngineParams &ep, const HostedIO &io, long long c) {
    const bbb_problem_spec &P = ep.spec;
    const bbb_buffers &B = ep.buf;
    int words = io.hrw;
    for (int s = 0; s < P.n_spaces; ++s)
        words += k_ref(B, s, (int)B.sel[s][c], c) * (P.spaces[s].ndim + P.spaces[s].n_params);
    return words;
}

}  // namespace bbb
