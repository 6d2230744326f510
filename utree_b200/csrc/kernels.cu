// kernels.cu -- sm_100a device side of the SEARCH_GG hot path + the thin C ABI
// over it (utb_db_*, utb_batch_*, stage-level entry points).
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   binix   : the CTR prefix index exactly as on disk, (2^24+1) x 4 B
//             (8 B when numNodes >= 2^32-1)                  itree.c:756-759
//   recs    : the CTR record blob exactly as on disk, numNodes x SZ bytes,
//             SZ = 5-byte suffix + IXTYPE id, + 32 B slack   itree.c:766
//             (irregular CTRs only; regular ones are re-laid out once at upload
//             into key / aux words, see DevDB, and get a Bloom filter)
//   labels  : blob of NUL-terminated strings, off[], rank[], by_rank[]
// Per batch (one stream slot):
//   raw     : the FASTA bytes as read from the file (headers included)
//   nl      : positions of the newlines (device-side framing)
//   seq_off/seq_len/name_off/name_len/grp_off : where each read's lines sit, and
//             the first 32-base group it owns in the packed "super-sequence"
//   pk/bad  : 2-bit codes (u64 per 32 bases, first base most significant, the
//             k-mer word order of itree.c:924) and a bad-base bit mask.  Every
//             read is padded to a multiple of 32 positions with >= 1 bad
//             position, so a 32-mer window can never straddle two reads.
//   hits    : one u32 per (position, strand): label id; valid where hitmap is set
//   results : one utb_result per read;  text : the output lines
//
// Kernels: nl_count / nl_index / frame_parse (XT_INITIATE_WS, itree.c:860-901),
// pack_kernel (XT_WORD_SEARCH's 2-bit packing, itree.c:919-926), filter_kernel or
// partition_kernel + probe_kernel (membership pre-filter, one fetch per position
// for both strands) and queue_lookup_kernel (XT_getIX32 + xtSuffixBS on the
// survivors, itree.c:699-730), lookup_kernel (the same without pre-filter, or the
// reference's exact probe sequence), vote_thread / vote_warp / vote_block (full
// aufbau vote, itree.c:1028-1098), fmt_* (the fprintf lines, itree.c:1032-1096).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "utb_internal.h"

namespace cg = cooperative_groups;
#define SUFMASK 0xFFFFFFFFFFull
#define HIT_MISS 0xFFFFFFFFu   // looked up, not found (BAD_IX widened)
#define HIT_NOWIN 0xFFFFFFFEu  // no valid 32-mer window here: no lookup made

#define CK(call)                                                                 \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) {                                                 \
            utb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_),   \
                          __FILE__, __LINE__, cudaGetErrorString(e_));           \
            return UTB_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

// ---------------------------------------------------------------------------
// device-side database view (passed by value to kernels)
// ---------------------------------------------------------------------------
struct DevDB {
    const uint32_t *binix32;   // one of binix32 / binix64 is non-null
    const uint64_t *binix64;
    const uint8_t *recs;
    uint64_t num_nodes;
    uint32_t sz;               // 7 or 9
    uint32_t ix_bytes;         // 2 or 4
    uint32_t max_ix;
    const char *blob;
    const uint32_t *off;
    const uint32_t *rank;
    const uint32_t *by_rank;
    uint32_t quirk_bin;        // bucket whose record 0 is the folded foreign record, else 0xFFFFFFFF
    // regular CTRs only: structure-of-arrays copy of the records (recs is then freed)
    const uint32_t *kv;        // IXTYPE uint16_t: blocks of 16 records = one 128-byte line, 16 keys (suffix >> 8) then their
                               // 16 aux words (suffix & 0xFF | id << 8): the line a key window misses on also holds its aux
    const uint32_t *keys;      // IXTYPE uint32_t: suffix >> 8 ...
    const uint64_t *aux64;     // ... and suffix & 0xFF | id << 8 in a second array
    // membership pre-filter over all record words (register-blocked Bloom, 16-byte blocks)
    const uint4 *bloom;
    uint64_t bloom_blocks;
};

struct utb_db {
    int device;
    DevDB d;
    int regular;               // every bucket strictly sorted (modulo quirk_bin): interpolation search is exact
    int use_interp;            // lookup kernel variant in use
    void *binix, *recs, *blob, *off, *rank, *by_rank, *keys, *aux, *bloom;
    int bloom_mode;            // 0 off, 1 always, 2 auto (on while the observed hit rate is low)
    volatile double ema_hit_rate;   // of the batches seen so far (written by the formatter thread, read at submit)
    size_t max_label;          // longest label, bytes
    volatile double text_per_read;  // running estimate of output bytes per read (sizes the speculative text D2H)
    uint64_t hbm_bytes;
    int l2_window;             // persisting-L2 window over binix configured
    size_t l2_window_bytes;
    float l2_hit_ratio;
};

// ---------------------------------------------------------------------------
// pack: raw bytes -> 2-bit groups + bad mask
// ---------------------------------------------------------------------------
// One thread per 32-base group.  The 32 bytes are fetched with nine aligned
// 32-bit loads (neighbouring threads share cache lines, so the raw bytes move
// once from L2), classified four at a time with byte-SIMD compares.
__device__ __forceinline__ uint32_t classify4(uint32_t w, uint32_t &badbits) {
    uint32_t u = w | 0x20202020u;                      // fold case (itree.c:114-117)
    uint32_t isA = __vcmpeq4(u, 0x61616161u), isC = __vcmpeq4(u, 0x63636363u);
    uint32_t isG = __vcmpeq4(u, 0x67676767u), isT = __vcmpeq4(u, 0x74747474u);
    uint32_t code = (isC & 0x01010101u) | (isG & 0x02020202u) | (isT & 0x03030303u);
    uint32_t inval = ~(isA | isC | isG | isT) & 0x01010101u;
    badbits = (inval * 0x01020408u) >> 24;             // bit j = byte j is not ACGTacgt
    // first base most significant
    return ((code & 3u) << 6) | (((code >> 8) & 3u) << 4) | (((code >> 16) & 3u) << 2) | ((code >> 24) & 3u);
}

__global__ void __launch_bounds__(256)
pack_kernel(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ seq_off,
            const uint32_t *__restrict__ seq_len, const uint32_t *__restrict__ grp_off,
            uint32_t n_reads, uint32_t n_groups, const uint32_t *__restrict__ n_groups_dev,
            uint64_t *__restrict__ pk, uint32_t *__restrict__ bad) {
    if (n_groups_dev) n_groups = *n_groups_dev;                    // device-side framing: the host only knows an upper bound
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (g == n_groups) { pk[g] = 0; bad[g] = 0xFFFFFFFFu; return; }   // guard group
    // read owning group g: largest r with grp_off[r] <= g
    uint32_t lo = 0, hi = n_reads;                     // invariant: grp_off[lo] <= g < grp_off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(grp_off + mid) <= g) lo = mid; else hi = mid;
    }
    uint32_t k = g - __ldg(grp_off + lo);
    uint32_t len = __ldg(seq_len + lo);
    uint32_t b0 = k * 32u;
    uint32_t nv = len > b0 ? min(32u, len - b0) : 0u;  // real bases in this group
    uint64_t word = 0;
    uint32_t badm = 0;
    if (nv) {
        uint64_t a = __ldg(seq_off + lo) + b0;
        const uint32_t *p = reinterpret_cast<const uint32_t *>(raw + (a & ~3ull));
        uint32_t sh = (uint32_t)(a & 3u) * 8u;
        uint32_t nw = (nv + 3u) >> 2;                  // 4-byte words that hold real bases
        uint32_t prev = __ldg(p);
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            uint32_t cur = (j < nw) ? __ldg(p + j + 1) : 0u;   // p[j+1] only if word j is needed
            uint32_t w = __funnelshift_r(prev, cur, sh);
            prev = cur;
            uint32_t bb;
            uint32_t c = classify4(w, bb);
            word = (word << 8) | c;
            badm |= bb << (4u * j);
        }
    }
    if (nv < 32u) badm |= 0xFFFFFFFFu << nv;           // padding positions are bad
    pk[g] = word;
    bad[g] = badm;
}

// ---------------------------------------------------------------------------
// window extraction (shared by lookup and the stage-level test kernel)
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool window_at(const uint64_t *__restrict__ pk, const uint32_t *__restrict__ bad,
                                          uint32_t pos, uint64_t &w) {
    uint32_t g = pos >> 5, o = pos & 31u;
    uint64_t hi = pk[g], lo = pk[g + 1];                          // const __restrict__ global pointers still compile to LDG.CONSTANT
    uint32_t bh = bad[g], bl = bad[g + 1];
    uint32_t wb = o ? ((bh >> o) | (bl << (32u - o))) : bh;
    w = o ? ((hi << (2u * o)) | (lo >> (64u - 2u * o))) : hi;
    return wb == 0;
}

// reverse complement of a 32-mer word: reverse the 2-bit fields of ~w
__device__ __forceinline__ uint64_t revcomp_word(uint64_t w) {
    uint64_t x = __brevll(~w);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// ---------------------------------------------------------------------------
// lookup: XT_getIX32 + xtSuffixBS, the reference's probe sequence verbatim
// ---------------------------------------------------------------------------
// Records are byte-packed (SZ = 7 or 9) so a suffix starts at any byte; its 5
// bytes always fall inside two consecutive aligned 32-bit words, which are the
// only bytes fetched (no sector is touched that the reference would not touch).
__device__ __forceinline__ uint64_t load_suffix(const uint8_t *__restrict__ recs, uint64_t byte_addr) {
    const uint32_t *p = reinterpret_cast<const uint32_t *>(recs + (byte_addr & ~3ull));
    uint32_t w0 = __ldg(p), w1 = __ldg(p + 1);
    uint32_t sh = (uint32_t)(byte_addr & 3u) * 8u;
    uint64_t v = ((uint64_t)w1 << 32) | w0;
    return (v >> sh) & SUFMASK;
}

__device__ __forceinline__ uint32_t load_ix(const DevDB &db, uint64_t rec) {
    const uint8_t *r = db.recs + rec * db.sz + 5;
    uint32_t v = (uint32_t)__ldg(r) | ((uint32_t)__ldg(r + 1) << 8);
    if (db.ix_bytes == 4) v |= ((uint32_t)__ldg(r + 2) << 16) | ((uint32_t)__ldg(r + 3) << 24);
    return v;
}

struct Probe {            // state of one in-flight xtSuffixBS
    uint64_t pos, size, suf;
    bool live;            // bucket non-empty
};

__device__ __forceinline__ void probe_begin(const DevDB &db, uint64_t word, Probe &q) {
    uint64_t p = word >> 40;
    uint64_t a, b;
    if (db.binix32) { a = __ldg(db.binix32 + p); b = __ldg(db.binix32 + p + 1); }
    else { a = __ldg(db.binix64 + p); b = __ldg(db.binix64 + p + 1); }
    q.suf = word & SUFMASK;
    q.live = a < b;                                    // itree.c:726
    // a CTR whose index points past the blob would make the reference read
    // out of bounds; clamp so the device never does
    if (b > db.num_nodes) q.live = false;
    q.pos = a;
    q.size = q.live ? b - a - 1 : 0;                   // itree.c:728
}

template <int N>
__device__ __forceinline__ void probe_run(const DevDB &db, Probe (&q)[N]) {
    // Lock-step binary searches: every round issues the N independent loads
    // first, then resolves the N compares (itree.c:701-705).
    for (;;) {
        bool any = false;
#pragma unroll
        for (int i = 0; i < N; ++i) any |= q[i].size != 0;
        if (!any) break;
        uint64_t v[N], h[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            h[i] = q[i].size >> 1;
            v[i] = q[i].size ? load_suffix(db.recs, (q[i].pos + h[i] + 1) * db.sz) : 0;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (q[i].size) {
                if (v[i] <= q[i].suf) { q[i].pos += h[i] + 1; q[i].size -= h[i] + 1; }
                else q[i].size = h[i];
            }
        }
    }
}

__device__ __forceinline__ uint32_t probe_end(const DevDB &db, const Probe &q) {
    if (!q.live) return HIT_MISS;
    if (load_suffix(db.recs, q.pos * db.sz) != q.suf) return HIT_MISS;   // itree.c:706
    uint32_t ix = load_ix(db, q.pos);
    return ix < db.max_ix ? ix : HIT_MISS;                                // itree.c:929
}


// ---- key-window search (regular CTRs only) -----------------------------------
// What bounds the exact-probe kernel is not bytes but the chain of dependent
// loads a warp has to wait for (profiles/r01_*: DRAM ~35 % busy, all warps
// resident, ~8 serialized round trips per lookup).  For a regular CTR (every
// bucket strictly sorted -- verified on the device at upload) the records are
// therefore re-laid out once, at load time, into 32-bit words:
//     key[i] = suffix >> 8          aux[i] = suffix & 0xFF | id << 8
// (uint16_t labels: blocks of 16 keys followed by their 16 aux words, one
// 128-byte line per block; uint32_t labels: two arrays).
// Suffixes are close to uniform inside a bucket, so record a + n*s/2^40 is
// within a few slots of the answer: ONE aligned 32-byte sector around it
// holds 8 candidate keys, which are ranked branch-free in registers.  A
// lookup is then index -> key line -> (hits only) aux: 2-3 dependent steps
// instead of ~8, and ~1.3 random DRAM fetches instead of ~5.  On a strictly
// sorted bucket any exact membership test returns what xtSuffixBS returns, so
// the result is identical; CTRs that are not regular keep the exact kernel.
enum { FW_MISS = 0, FW_FOUND = 1, FW_LEFT = 2, FW_RIGHT = 3, FW_SLOW = 4 };
#define KW 8u                   // keys per window = one 32-byte sector
struct FastProbe {
    uint64_t a, b;      // bucket [a,b), quirk-adjusted
    uint64_t ws;        // first key of the current window (multiple of KW)
    uint64_t pos;       // FW_FOUND: index of the first key equal to t in the window
    uint32_t cnt;       // FW_FOUND: length of the run of equal keys (records that differ only in the suffix's low byte)
    uint32_t t, lo8;    // target key and the suffix's low byte
    int st;
};
__device__ __forceinline__ void fast_begin(const DevDB &db, uint64_t word, FastProbe &q) {
    uint64_t p = word >> 40;
    uint64_t a, b;
    if (db.binix32) { a = __ldg(db.binix32 + p); b = __ldg(db.binix32 + p + 1); }
    else { a = __ldg(db.binix64 + p); b = __ldg(db.binix64 + p + 1); }
    uint64_t suf = word & SUFMASK;
    q.t = (uint32_t)(suf >> 8); q.lo8 = (uint32_t)(suf & 0xFFu);
    bool live = a < b && b <= db.num_nodes;                        // itree.c:726 (+ bounds guard)
    if (live && (uint32_t)p == db.quirk_bin) { a += 1; live = a < b; }   // record 0 is unreachable for xtSuffixBS
    q.a = a; q.b = b; q.pos = 0; q.cnt = 0;
    q.st = live ? FW_LEFT : FW_MISS;                               // any non-final state: "window pending"
    q.ws = live ? ((a + __umul64hi(suf << 24, b - a)) & ~(uint64_t)(KW - 1)) : 0;   // a + floor(n*s/2^40) < b
}
// One sector of keys around the estimate.  The miss fills the whole 128-byte
// line in L2, so the neighbouring sectors a later step may need are L2 hits;
// asking for them up front would cost as much as further misses
// (profiles/r01_membench_ncu.txt).
__device__ __forceinline__ uint64_t kv_index(uint64_t i) { return ((i >> 4) << 5) | (i & 15u); }   // u32 index of key i in db.kv
__device__ __forceinline__ uint32_t key_at(const DevDB &db, uint64_t i) { return db.kv ? __ldg(db.kv + kv_index(i)) : __ldg(db.keys + i); }
__device__ __forceinline__ void window_fetch(const DevDB &db, const FastProbe &q, uint4 &k0, uint4 &k1) {
    const uint4 *p = reinterpret_cast<const uint4 *>(db.kv ? db.kv + kv_index(q.ws) : db.keys + q.ws);   // ws is a multiple of 8
    k0 = __ldg(p); k1 = __ldg(p + 1);
}
__device__ __forceinline__ void window_rank(FastProbe &q, const uint4 &k0, const uint4 &k1) {
    const uint32_t kk[KW] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const uint64_t wa = q.ws > q.a ? q.ws : q.a, wb = q.ws + KW < q.b ? q.ws + KW : q.b;
    const uint32_t j0 = (uint32_t)(wa - q.ws), j1 = (uint32_t)(wb - q.ws);   // in-bucket slots [j0, j1)
    uint32_t lt = 0, eq = 0;
#pragma unroll
    for (uint32_t j = 0; j < KW; ++j) {
        bool in = j >= j0 && j < j1;
        lt += in && kk[j] < q.t;
        eq += in && kk[j] == q.t;
    }
    const bool left_open = wa > q.a, right_open = wb < q.b;
    if (eq == 0) {
        if (lt == 0 && left_open) { q.st = FW_LEFT; q.ws -= KW; }             // every key here is larger
        else if (lt == j1 - j0 && right_open) { q.st = FW_RIGHT; q.ws += KW; } // every key here is smaller
        else q.st = FW_MISS;
    } else { q.st = FW_FOUND; q.pos = wa + lt; q.cnt = eq; }       // eq > 1: k-mers of related genomes that differ in the last 4 bases
}
__device__ __forceinline__ uint64_t load_aux(const DevDB &db, uint64_t i) {
    return db.kv ? (uint64_t)__ldg(db.kv + kv_index(i) + 16u) : __ldg(db.aux64 + i);
}
// rare: exact lower bound over the whole bucket on the full 40-bit suffix
__device__ __noinline__ uint32_t fast_slow(const DevDB &db, uint64_t a, uint64_t b, uint32_t t, uint32_t lo8) {
    const uint64_t s = ((uint64_t)t << 8) | lo8;
    uint64_t lo = a, hi = b;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        uint64_t c = ((uint64_t)key_at(db, mid) << 8) | (load_aux(db, mid) & 0xFFu);
        if (c < s) lo = mid + 1; else hi = mid;
    }
    if (lo >= b) return HIT_MISS;
    uint64_t ax = load_aux(db, lo);
    if ((((uint64_t)key_at(db, lo) << 8) | (ax & 0xFFu)) != s) return HIT_MISS;
    uint32_t ix = (uint32_t)(ax >> 8);
    return ix < db.max_ix ? ix : HIT_MISS;
}
#define FW_MAX_STEPS 4          // sectors inspected before giving up on the estimate
// sect: 32-byte sectors this lookup had to touch (index pair, key windows, aux)
template <int N>
__device__ __forceinline__ void fast_lookup(const DevDB &db, const uint64_t (&w)[N], uint32_t (&r)[N], uint32_t *sect = nullptr) {
    FastProbe q[N];
    uint32_t ns = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) fast_begin(db, w[i], q[i]);
    int dir[N];
    bool act[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ns += ((w[i] >> 40) & 7u) == 7u ? 2u : 1u;                 // BinIx[p], BinIx[p+1]: 4-byte entries, 8 per sector
        dir[i] = 0; act[i] = q[i].st != FW_MISS;                   // empty bucket: nothing to search
    }
    // the N searches advance in lock-step: all key sectors of a step are requested before any is ranked
    for (int step = 0; step < FW_MAX_STEPS; ++step) {
        bool any = false;
#pragma unroll
        for (int i = 0; i < N; ++i) any |= act[i];
        if (!any) break;
        uint4 k0[N], k1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) if (act[i]) { ++ns; window_fetch(db, q[i], k0[i], k1[i]); }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (!act[i]) continue;
            window_rank(q[i], k0[i], k1[i]);
            if (q[i].st != FW_LEFT && q[i].st != FW_RIGHT) act[i] = false;
            else if (dir[i] && q[i].st != dir[i]) { q[i].st = FW_MISS; act[i] = false; }   // turned around: between two adjacent windows
            else dir[i] = q[i].st;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) if (q[i].st == FW_LEFT || q[i].st == FW_RIGHT) q[i].st = FW_SLOW;   // estimate off by several sectors
    // matches fetch low byte + id of every record in the run of equal keys (one sector of aux); a neighbour key is
    // checked when the run touches an open window edge, because it could continue outside
#pragma unroll
    for (int i = 0; i < N; ++i) {
        r[i] = HIT_MISS;
        if (q[i].st == FW_FOUND) {
            ++ns;
            const uint64_t wa = q[i].ws > q[i].a ? q[i].ws : q[i].a, wb = q[i].ws + KW < q[i].b ? q[i].ws + KW : q[i].b;
            uint32_t nbl = ~q[i].t, nbr = ~q[i].t;
            if (q[i].pos == wa && wa > q[i].a) nbl = key_at(db, q[i].pos - 1);
            if (q[i].pos + q[i].cnt == wb && wb < q[i].b) nbr = key_at(db, q[i].pos + q[i].cnt);
            uint64_t ax[KW];
#pragma unroll
            for (uint32_t j = 0; j < KW; ++j) ax[j] = j < q[i].cnt ? load_aux(db, q[i].pos + j) : 0;
            if (nbl == q[i].t || nbr == q[i].t) q[i].st = FW_SLOW; // the run of equal keys crosses the window
            else {
#pragma unroll
                for (uint32_t j = 0; j < KW; ++j)
                    if (j < q[i].cnt && (uint32_t)(ax[j] & 0xFFu) == q[i].lo8) { const uint32_t ix = (uint32_t)(ax[j] >> 8); r[i] = ix < db.max_ix ? ix : HIT_MISS; }
            }
        }
        if (q[i].st == FW_SLOW) { r[i] = fast_slow(db, q[i].a, q[i].b, q[i].t, q[i].lo8); ns += 8; }
    }
    if (sect) *sect = ns;
}

// raw CTR records -> SoA (runs once at upload, only for regular CTRs)
__global__ void __launch_bounds__(256)
relayout_kernel(DevDB db, uint32_t *__restrict__ kv, uint32_t *__restrict__ keys, uint64_t *__restrict__ aux64) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= db.num_nodes) return;
    uint64_t suf = load_suffix(db.recs, i * db.sz);
    uint32_t ix = load_ix(db, i);
    if (kv) { kv[kv_index(i)] = (uint32_t)(suf >> 8); kv[kv_index(i) + 16u] = (uint32_t)(suf & 0xFFu) | (ix << 8); }
    else { keys[i] = (uint32_t)(suf >> 8); aux64[i] = (suf & 0xFFu) | ((uint64_t)ix << 8); }
}

// Load-time check of the invariant the interpolation search relies on.
// out[0] buckets with a disorder past their first pair (or an index beyond the blob)
// out[1] buckets whose only disorder is record0 >= record1   out[2] smallest such bin
// out[3] first non-empty bin
__global__ void __launch_bounds__(256)
verify_kernel(DevDB db, unsigned long long *__restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= UTB_NUMBINS - 1) return;
    uint64_t a, b;
    if (db.binix32) { a = db.binix32[p]; b = db.binix32[p + 1]; }
    else { a = db.binix64[p]; b = db.binix64[p + 1]; }
    if (a >= b) return;                                            // empty for XT_getIX32 (itree.c:726)
    if (b > db.num_nodes) { atomicAdd(out + 0, 1ull); return; }
    atomicMin(out + 3, (unsigned long long)p);
    uint64_t prev = load_suffix(db.recs, a * db.sz);
    bool first_bad = false, rest_bad = false;
    for (uint64_t i = a + 1; i < b; ++i) {
        uint64_t cur = load_suffix(db.recs, i * db.sz);
        if (cur <= prev) { if (i == a + 1) first_bad = true; else rest_bad = true; }
        prev = cur;
    }
    if (rest_bad) atomicAdd(out + 0, 1ull);
    else if (first_bad) { atomicAdd(out + 1, 1ull); atomicMin(out + 2, (unsigned long long)p); }
}

// ---- membership pre-filter ----------------------------------------------------
// Most 32-mers of a read are not in a sampled tree (complevel 2 keeps 1/16 of
// the k-mers), and every miss costs the exact path ~2.6 random DRAM fetches.
// A register-blocked Bloom filter over the record words, built on the device
// at upload, answers "certainly absent" with ONE 16-byte load: block =
// 4 x u32, two bits per word, all eight tests on registers.  No false negatives,
// so a filtered miss is a miss of XT_getIX32 too; positives take the exact path.
//
// The BLOCK of a word x is chosen by its strand-neutral form c = min(x, revcomp(x)),
// the eight BITS by x itself.  A read position looks up x and revcomp(x)
// (itree.c:887-898 appends the reverse complement, so both strands of every
// window are searched): both share c, hence both tests are served by the same
// 16-byte load -- one scattered fetch per POSITION instead of one per lookup.
// Each record still sets eight bits of one block, so the false-positive rate is
// that of the plain blocked filter (~0.1 % at 16 bits per record).
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull;
    x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull;
    return x ^ (x >> 33);
}
// hc = mix64(min(x, revcomp(x))); the mask hash folds the word back in, so x and revcomp(x) get unrelated bits
__device__ __forceinline__ uint4 bloom_masks(uint64_t hc, uint64_t x) {
    const uint64_t g = ((hc ^ x) * 0x9E3779B97F4A7C15ull) >> 24;   // 8 x 5 bits: two bits in each of the 4 words
    return make_uint4((1u << (g & 31)) | (1u << ((g >> 5) & 31)), (1u << ((g >> 10) & 31)) | (1u << ((g >> 15) & 31)),
                      (1u << ((g >> 20) & 31)) | (1u << ((g >> 25) & 31)), (1u << ((g >> 30) & 31)) | (1u << ((g >> 35) & 31)));
}
__device__ __forceinline__ bool bloom_test(const uint4 &v, const uint4 &m) {
    return ((v.x & m.x) == m.x) & ((v.y & m.y) == m.y) & ((v.z & m.z) == m.z) & ((v.w & m.w) == m.w);
}
__device__ __forceinline__ uint64_t bloom_block_hash(uint64_t x, uint64_t rc) { return mix64(x < rc ? x : rc); }
__device__ __forceinline__ bool bloom_maybe(const DevDB &db, uint64_t word) {
    const uint64_t hc = bloom_block_hash(word, revcomp_word(word));
    const uint4 v = __ldg(db.bloom + __umul64hi(hc, db.bloom_blocks));   // ONE request for ONE line (profiles/r01_membench*)
    return bloom_test(v, bloom_masks(hc, word));
}
// one thread per prefix bin: inserts the words of the bin's records
__global__ void __launch_bounds__(256)
bloom_build_kernel(DevDB db, uint32_t *__restrict__ bloom) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= UTB_NUMBINS - 1) return;
    uint64_t a, b;
    if (db.binix32) { a = db.binix32[p]; b = db.binix32[p + 1]; }
    else { a = db.binix64[p]; b = db.binix64[p + 1]; }
    if (a >= b || b > db.num_nodes) return;
    if (p == db.quirk_bin) ++a;                                    // the folded record is unreachable anyway
    for (uint64_t i = a; i < b; ++i) {
        const uint64_t word = ((uint64_t)p << 40) | ((uint64_t)key_at(db, i) << 8) | (load_aux(db, i) & 0xFFu);
        const uint64_t hc = bloom_block_hash(word, revcomp_word(word));
        uint32_t *blk = bloom + 4 * __umul64hi(hc, db.bloom_blocks);
        const uint4 m = bloom_masks(hc, word);
        atomicOr(blk + 0, m.x); atomicOr(blk + 1, m.y); atomicOr(blk + 2, m.z); atomicOr(blk + 3, m.w);
    }
}

// One thread per (position, strand) of the packed super-sequence: lanes 2k and
// 2k+1 look up the forward and the reverse-complement word of window k, so the
// hit slots of a warp are 32 consecutive u32.  No block-level synchronisation:
// a warp that is done leaves (the profile of the first version showed a third
// of the resident warps parked on a bookkeeping barrier); lookups and hits are
// counted per warp into COUNTER_SLOTS spread counters that the host sums.
#define COUNTER_SLOTS 1024
template <int NSTR, bool FAST, bool BLOOM>
__global__ void __launch_bounds__(256, 6)
lookup_kernel(DevDB db, const uint64_t *__restrict__ pk, const uint32_t *__restrict__ bad,
              uint32_t n_pos, const uint32_t *__restrict__ n_groups_dev, uint32_t *__restrict__ hits, unsigned long long *__restrict__ counters) {
    if (n_groups_dev) n_pos = *n_groups_dev * 32u;
    const uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pos = (uint32_t)(NSTR == 2 ? slot >> 1 : slot);
    uint64_t w = 0;
    const bool valid = pos < n_pos && window_at(pk, bad, pos, w);
    uint32_t r = HIT_NOWIN;
    if (valid) {
        if (NSTR == 2 && (slot & 1)) w = revcomp_word(w);
        if (BLOOM && !bloom_maybe(db, w)) r = HIT_MISS;
        else if (FAST) {
            uint64_t ww[1] = {w};
            uint32_t rr[1];
            fast_lookup<1>(db, ww, rr);
            r = rr[0];
        } else {
            Probe q[1];
            probe_begin(db, w, q[0]);
            probe_run<1>(db, q);
            r = probe_end(db, q[0]);
        }
    }
    if (pos < n_pos) hits[slot] = r;
    const uint32_t nv = __popc(__ballot_sync(0xFFFFFFFFu, valid));
    const uint32_t nh = __popc(__ballot_sync(0xFFFFFFFFu, r < HIT_NOWIN));
    if ((threadIdx.x & 31u) == 0) {
        const uint32_t c = blockIdx.x & (COUNTER_SLOTS - 1);
        if (nv) atomicAdd(counters + c, (unsigned long long)nv);
        if (nh) atomicAdd(counters + COUNTER_SLOTS + c, (unsigned long long)nh);
    }
}

// ---- two-phase lookup (pre-filter on) -------------------------------------------
// In one fused kernel every warp waits for its slowest lane, i.e. for the one
// or two lanes per warp whose word passes the filter and walks the whole
// index -> key line -> aux chain, while the other ~30 lanes idle: the kernel
// is bound by warp latency, not by DRAM.  So the work is split into two dense
// passes: phase A tests every window against the filter (one fetch per POSITION
// serves both strands, no chain) and appends the survivors to a queue with a
// warp-aggregated atomic; phase B runs the exact search over the queue with
// every lane busy.
__device__ __noinline__ uint32_t fast_lookup_cold(const DevDB &db, uint64_t w) {
    uint64_t ww[1] = {w};
    uint32_t rr[1];
    fast_lookup<1>(db, ww, rr);
    return rr[0];
}
#define Q_CHUNK 128u              // queue slots a warp reserves per atomic (>= 64: two ballots per step)
#ifndef FILT_ILP
#define FILT_ILP 2                // filter probes in flight per lane
#endif
#define Q_INVALID 0xFFFFFFFFu
#ifndef FILT_MINB
#define FILT_MINB 4               // CTAs per SM the register budget is sized for
#endif
// A warp's private window into the survivor queue: the global counter sees one
// atomic per Q_CHUNK survivors instead of one per warp-step.
struct WarpQueue {
    uint64_t base;
    uint32_t used;
    bool have;
};
__device__ __forceinline__ void wq_init(WarpQueue &q) { q.base = 0; q.used = Q_CHUNK; q.have = false; }
__device__ __forceinline__ void wq_retire(const WarpQueue &q, uint32_t lane, uint32_t *__restrict__ q_slots) {
    if (q.have) for (uint32_t j = q.used + lane; j < Q_CHUNK; j += 32) q_slots[q.base + j] = Q_INVALID;   // pad the tail
}
// make room for c more entries (warp-uniform call)
__device__ __forceinline__ void wq_reserve(WarpQueue &q, uint32_t c, uint32_t lane, uint32_t *__restrict__ q_slots,
                                           unsigned long long *__restrict__ q_count, uint64_t q_cap) {
    if (q.used + c <= Q_CHUNK) return;
    wq_retire(q, lane, q_slots);
    unsigned long long nb = 0;
    if (lane == 0) nb = atomicAdd(q_count, (unsigned long long)Q_CHUNK);
    q.base = __shfl_sync(0xFFFFFFFFu, nb, 0);
    q.used = 0;
    q.have = q.base + Q_CHUNK <= q_cap;                            // beyond capacity: the caller resolves inline
}
// queue overflow / region overflow: the exact search right here (rare)
__device__ __noinline__ uint32_t resolve_inline(const DevDB &db, uint64_t w, uint32_t slot, uint32_t *__restrict__ hits,
                                                uint32_t *__restrict__ hitmap) {
    const uint32_t r = fast_lookup_cold(db, w);
    if (r == HIT_MISS) return 0;
    hits[slot] = r; atomicOr(hitmap + (slot >> 5), 1u << (slot & 31u));
    return 1;
}
// Tests one position (both strands when NSTR == 2) against its filter block v and appends the survivors.
// Warp-uniform call; `valid` lanes hold a window w with reverse complement rc and block hash hc.
template <int NSTR>
__device__ __forceinline__ uint32_t filter_emit(const DevDB &db, WarpQueue &wq, uint32_t lane, bool valid, uint64_t w, uint64_t rc,
                                                uint64_t hc, const uint4 &v, uint32_t pos,
                                                uint64_t *__restrict__ q_words, uint32_t *__restrict__ q_slots,
                                                unsigned long long *__restrict__ q_count, uint64_t q_cap,
                                                uint32_t *__restrict__ hits, uint32_t *__restrict__ hitmap) {
    const bool passF = valid && bloom_test(v, bloom_masks(hc, w));
    const bool passR = NSTR == 2 && valid && bloom_test(v, bloom_masks(hc, rc));
    const uint32_t bF = __ballot_sync(0xFFFFFFFFu, passF);
    const uint32_t bR = NSTR == 2 ? __ballot_sync(0xFFFFFFFFu, passR) : 0u;
    const uint32_t cF = __popc(bF), c = cF + __popc(bR);
    uint32_t nh = 0;
    if (!c) return 0;
    wq_reserve(wq, c, lane, q_slots, q_count, q_cap);
    const uint32_t lt = (1u << lane) - 1u;
    if (wq.have) {
        const uint64_t at = wq.base + wq.used;
        if (passF) { const uint64_t i = at + __popc(bF & lt); q_words[i] = w; q_slots[i] = pos * NSTR; }
        if (passR) { const uint64_t i = at + cF + __popc(bR & lt); q_words[i] = rc; q_slots[i] = pos * NSTR + 1u; }
    } else {
        if (passF) nh += resolve_inline(db, w, pos * NSTR, hits, hitmap);
        if (passR) nh += resolve_inline(db, rc, pos * NSTR + 1u, hits, hitmap);
    }
    wq.used += c;
    return nh;
}

// Direct filter pass (batches too small for the partitioned pass): persistent warps, one lane per
// position, FILT_ILP positions of a lane in flight together.
template <int NSTR>
__global__ void __launch_bounds__(256, FILT_MINB)
filter_kernel(DevDB db, const uint64_t *__restrict__ pk, const uint32_t *__restrict__ bad, uint32_t n_pos,
              const uint32_t *__restrict__ n_groups_dev,
              uint32_t *__restrict__ hits, unsigned long long *__restrict__ counters,
              uint64_t *__restrict__ q_words, uint32_t *__restrict__ q_slots, unsigned long long *__restrict__ q_count,
              uint64_t q_cap, uint32_t *__restrict__ hitmap) {
    if (n_groups_dev) n_pos = *n_groups_dev * 32u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    WarpQueue wq;
    wq_init(wq);
    uint32_t nv = 0, nh = 0;
    // every warp takes FILT_ILP x 32 consecutive positions per round (n_pos is a multiple of 32)
    for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * FILT_ILP; base < n_pos; base += stride * FILT_ILP) {
        uint64_t w[FILT_ILP], hc[FILT_ILP];
        uint4 v[FILT_ILP];
        bool valid[FILT_ILP];
#pragma unroll
        for (int u = 0; u < FILT_ILP; ++u) {
            const uint64_t pos = base + 32u * u + lane;
            w[u] = 0;
            valid[u] = pos < n_pos && window_at(pk, bad, (uint32_t)pos, w[u]);
            hc[u] = bloom_block_hash(w[u], revcomp_word(w[u]));
            v[u] = make_uint4(0, 0, 0, 0);
            if (valid[u]) v[u] = __ldg(db.bloom + __umul64hi(hc[u], db.bloom_blocks));
        }
#pragma unroll
        for (int u = 0; u < FILT_ILP; ++u) {
            const uint32_t pos = (uint32_t)(base + 32u * u + lane);
            // the reverse complement is recomputed rather than kept live across the loads (registers = probes in flight)
            nh += filter_emit<NSTR>(db, wq, lane, valid[u], w[u], revcomp_word(w[u]), hc[u], v[u], pos, q_words, q_slots, q_count, q_cap, hits, hitmap);
            nv += valid[u] ? NSTR : 0;                             // hits[] is only valid where hitmap is set: nothing to store for misses
        }
    }
    wq_retire(wq, lane, q_slots);
    for (int o = 16; o; o >>= 1) { nv += __shfl_xor_sync(0xFFFFFFFFu, nv, o); nh += __shfl_xor_sync(0xFFFFFFFFu, nh, o); }
    if (lane == 0) {
        const uint32_t cs = (blockIdx.x * 8 + (threadIdx.x >> 5)) & (COUNTER_SLOTS - 1);
        if (nv) atomicAdd(counters + cs, (unsigned long long)nv);
        if (nh) atomicAdd(counters + COUNTER_SLOTS + cs, (unsigned long long)nh);
    }
}

// ---- partitioned filter pass ----------------------------------------------------------
// One random 128-byte DRAM line per probe is what bounds filter_kernel.  Here the
// positions of a batch are first bucketed by the top 6 bits of the filter's block
// hash (64 partitions = 64 contiguous ~34 MB slices of the filter): partition_kernel
// streams (forward word, position) records into per-partition arrays (12 B per
// position = 6 B per lookup with both strands, sequential), then probe_kernel sweeps
// the partitions in order with the whole grid, so the slice being probed stays
// resident in the 126 MB L2 and each of its lines is fetched from HBM once per batch
// instead of once per probe.
#define NPART 64u
#define PART_MIN_POS (256ull << 20)  // positions in a batch from which the partitioned pass is used (~1.6 M reads of 150 bp); below it the
                                     // direct filter wins: 64 grid barriers per batch and a cooperative launch that cannot overlap other slots
#define P_CTAS (148u * 4u)      // partitioner CTAs; each owns one output region per partition
#define P_TILE 4096u            // positions a CTA sorts in shared memory at a time
#define P_PPT (P_TILE / 256u)   // positions per thread and tile
#define P_SMEM (P_TILE * 8u + P_TILE * 4u)   // words in tile order + sorted order (local index | partition << 16)
// Partitioner: every CTA counting-sorts one 4096-position tile at a time and appends each partition's
// run to its own region [(c*NPART+p)*cap_cp, +cap_cp), so the global stores are contiguous runs (~47
// records) instead of one transaction per record (scattered 8-byte stores cost as much as scattered
// sector reads, profiles/).  The words stay in tile order in shared memory; only a 4-byte index is
// scattered, and the output pass gathers through it.  No global atomics; the fill of every region is
// stored at the end.  Hash partitions are uniform and every CTA sees the same share of the batch, so
// regions fill evenly (cap_cp carries 25 % head-room; records of an overflowing tile are resolved inline).
// MATCH (kept as a measured alternative, UTB_PART_ATOMS=0): the rank of a record inside its partition comes
// from warp match + a scanned (warp-row x partition) count table instead of one shared-memory atomic per
// record.  Shared atomics with a return value run at ~2 cycles per LANE on the SM's load/store unit
// (B300_MICROARCH "ATOMS spread-addr"), but MATCH.ANY over mostly distinct keys plus the table scan cost
// more: 16.2 ms against 11.5 ms per 10 M reads, so the atomic variant is the default.
#define P_ROWS (P_TILE / 32u)   // warp-rows of a tile
#define P_SMEM_MATCH (P_SMEM + P_ROWS * NPART * 2u)
template <int NSTR, bool MATCH>
__global__ void __launch_bounds__(256, MATCH ? 3 : 4)
partition_kernel(DevDB db, const uint64_t *__restrict__ pk, const uint32_t *__restrict__ bad, uint32_t n_pos,
                 const uint32_t *__restrict__ n_groups_dev, unsigned long long *__restrict__ counters,
                 uint64_t *__restrict__ p_words, uint32_t *__restrict__ p_pos, uint32_t *__restrict__ p_fill,
                 uint32_t cap_cp, uint32_t *__restrict__ hits, uint32_t *__restrict__ hitmap) {
    extern __shared__ __align__(16) unsigned char p_smem[];
    uint64_t *s_w = reinterpret_cast<uint64_t *>(p_smem);
    uint32_t *s_ord = reinterpret_cast<uint32_t *>(p_smem + P_TILE * 8u);
    uint16_t *s_cnt = reinterpret_cast<uint16_t *>(p_smem + P_SMEM);   // MATCH: [P_ROWS][NPART] group sizes, then their prefix over the rows
    __shared__ uint32_t s_q[4][NPART];                             // MATCH: per quarter of the rows: total, then base in the sorted order
    __shared__ uint32_t s_hist[NPART], s_start[NPART], s_cur[NPART];
    __shared__ uint64_t s_gbase[NPART];                            // region base + fill - start: global index = s_gbase[part] + sorted rank
    __shared__ uint32_t s_ovf;
    __shared__ uint64_t s_pk[P_TILE / 32u + 4u];                  // the tile's packed groups (+ the one after)
    __shared__ uint32_t s_bad[P_TILE / 32u + 4u];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (n_groups_dev) n_pos = *n_groups_dev * 32u;
    const uint32_t n_groups = n_pos >> 5;
    if (tid < NPART) s_cur[tid] = 0;
    uint32_t nv = 0, nh = 0;
    const uint64_t region0 = (uint64_t)blockIdx.x * NPART * cap_cp;
    // The packed stream of a tile is 1.5 KB: it is fetched one tile ahead into registers (the loads complete
    // under the three passes of the current tile) and parked in shared memory, so no pass waits on DRAM.
    uint64_t pf_pk = 0;
    uint32_t pf_bad = 0xFFFFFFFFu;
    {
        const uint64_t g = (uint64_t)blockIdx.x * (P_TILE / 32u) + tid;
        if (tid <= P_TILE / 32u && g <= n_groups) { pf_pk = __ldg(pk + g); pf_bad = __ldg(bad + g); }   // group n_groups is the guard
    }
    for (uint64_t tile = (uint64_t)blockIdx.x * P_TILE; tile < n_pos; tile += (uint64_t)gridDim.x * P_TILE) {
        if (tid < NPART) s_hist[tid] = 0;
        if (tid == 0) s_ovf = 0;
        if (tid <= P_TILE / 32u) { s_pk[tid] = pf_pk; s_bad[tid] = pf_bad; }
        if (MATCH) {
            uint4 *z = reinterpret_cast<uint4 *>(s_cnt);
#pragma unroll
            for (uint32_t i = 0; i < P_ROWS * NPART * 2u / 16u / 256u; ++i) z[i * 256u + tid] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        {
            const uint64_t g = (tile + (uint64_t)gridDim.x * P_TILE) / 32u + tid;
            pf_pk = 0; pf_bad = 0xFFFFFFFFu;
            if (tid <= P_TILE / 32u && g <= n_groups) { pf_pk = __ldg(pk + g); pf_bad = __ldg(bad + g); }
        }
        // pass A: window, partition id and rank inside the tile.  A warp covers 32 consecutive positions = one
        // group of the packed stream, so its pk/bad loads are warp-uniform.
        uint32_t meta[P_PPT];                                      // rank << 6 | part, or 0xFFFFFFFF
        const uint32_t tile_pos = (uint32_t)tile;
#pragma unroll
        for (uint32_t k = 0; k < P_PPT; ++k) {
            const uint32_t local = k * 256u + tid;
            uint64_t w = 0;
            meta[k] = 0xFFFFFFFFu;
            const bool valid = window_at(s_pk, s_bad, local, w);   // positions past the batch lie in all-bad groups
            uint32_t part = 64u + lane;                             // invalid lanes: a key of their own
            if (valid) part = (uint32_t)(bloom_block_hash(w, revcomp_word(w)) >> 58);
            if (MATCH) {
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, part);
                const uint32_t lrank = __popc(peers & ((1u << lane) - 1u));
                if (valid) {
                    if (lrank == 0) s_cnt[(k * 8u + (tid >> 5)) * NPART + part] = (uint16_t)__popc(peers);
                    meta[k] = (lrank << 6) | part;
                    nv += NSTR;
                }
            } else if (valid) {
                meta[k] = (atomicAdd(&s_hist[part], 1u) << 6) | part;
                nv += NSTR;
            }
            s_w[local] = w;
        }
        __syncthreads();
        if (MATCH) {
            // prefix of the group sizes over the warp-rows, per partition: thread = (quarter of the rows, partition)
            const uint32_t bin = tid & 63u, q = tid >> 6;
            uint32_t run = 0;
#pragma unroll 8
            for (uint32_t i = 0; i < P_ROWS / 4u; ++i) {
                uint16_t *c = s_cnt + (q * (P_ROWS / 4u) + i) * NPART + bin;
                const uint32_t v = *c;
                *c = (uint16_t)run;
                run += v;
            }
            s_q[q][bin] = run;
            __syncthreads();
            if (tid < NPART) s_hist[tid] = s_q[0][tid] + s_q[1][tid] + s_q[2][tid] + s_q[3][tid];
            __syncthreads();
        }
        if (tid < 32) {                                            // exclusive scan of the 64 bins by one warp
            const uint32_t h0 = s_hist[2 * tid], h1 = s_hist[2 * tid + 1];
            uint32_t x = h0 + h1;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o); if ((int)tid >= o) x += t; }
            const uint32_t st0 = x - h0 - h1, st1 = x - h1, c0 = s_cur[2 * tid], c1 = s_cur[2 * tid + 1];
            s_start[2 * tid] = st0; s_start[2 * tid + 1] = st1;
            s_gbase[2 * tid] = region0 + (uint64_t)(2 * tid) * cap_cp + c0 - st0;
            s_gbase[2 * tid + 1] = region0 + (uint64_t)(2 * tid + 1) * cap_cp + c1 - st1;
            if (c0 + h0 > cap_cp || c1 + h1 > cap_cp) s_ovf = 1;
            if (MATCH) {                                            // quarter totals -> bases in the sorted order
#pragma unroll
                for (uint32_t b2 = 0; b2 < 2; ++b2) {
                    uint32_t base = b2 ? st1 : st0;
#pragma unroll
                    for (uint32_t q = 0; q < 4; ++q) { const uint32_t t = s_q[q][2 * tid + b2]; s_q[q][2 * tid + b2] = base; base += t; }
                }
            }
        }
        __syncthreads();
        // pass B: the sorted order (4 bytes per record scattered; the words stay where they are)
#pragma unroll
        for (uint32_t k = 0; k < P_PPT; ++k) {
            const uint32_t m = meta[k];
            if (m == 0xFFFFFFFFu) continue;
            const uint32_t part = m & 63u;
            uint32_t dst;
            if (MATCH) { const uint32_t row = k * 8u + (tid >> 5); dst = s_q[row / (P_ROWS / 4u)][part] + s_cnt[row * NPART + part] + (m >> 6); }
            else dst = s_start[part] + (m >> 6);
            s_ord[dst] = (k * 256u + tid) | (part << 16);
        }
        __syncthreads();
        // pass C: contiguous runs out to the regions
        const uint32_t total = s_start[NPART - 1] + s_hist[NPART - 1];
        if (!s_ovf) {
            for (uint32_t e = tid; e < total; e += 256) {
                const uint32_t o = s_ord[e], local = o & 0xFFFFu;
                const uint64_t idx = s_gbase[o >> 16] + e;
                p_words[idx] = s_w[local]; p_pos[idx] = tile_pos + local;
            }
        } else {                                                   // some region is full: per-record bound check
            for (uint32_t e = tid; e < total; e += 256) {
                const uint32_t o = s_ord[e], local = o & 0xFFFFu, part = o >> 16;
                const uint32_t at = s_cur[part] + (e - s_start[part]);
                const uint64_t w = s_w[local];
                const uint32_t pos = tile_pos + local;
                if (at < cap_cp) { const uint64_t idx = s_gbase[part] + e; p_words[idx] = w; p_pos[idx] = pos; }
                else {
                    const uint64_t rc = revcomp_word(w);
                    if (bloom_maybe(db, w)) nh += resolve_inline(db, w, pos * NSTR, hits, hitmap);
                    if (NSTR == 2 && bloom_maybe(db, rc)) nh += resolve_inline(db, rc, pos * NSTR + 1u, hits, hitmap);
                }
            }
        }
        __syncthreads();
        if (tid < NPART) s_cur[tid] += s_hist[tid];
    }
    __syncthreads();
    if (tid < NPART) p_fill[blockIdx.x * NPART + tid] = s_cur[tid] < cap_cp ? s_cur[tid] : cap_cp;
    for (int o = 16; o; o >>= 1) { nv += __shfl_xor_sync(0xFFFFFFFFu, nv, o); nh += __shfl_xor_sync(0xFFFFFFFFu, nh, o); }
    if (lane == 0) {
        const uint32_t cs = (blockIdx.x * 8 + (tid >> 5)) & (COUNTER_SLOTS - 1);
        if (nv) atomicAdd(counters + cs, (unsigned long long)nv);
        if (nh) atomicAdd(counters + COUNTER_SLOTS + cs, (unsigned long long)nh);
    }
}

// Probe pass (cooperative launch): the whole grid sweeps partition 0, then 1, ... with a grid barrier in
// between, so exactly one ~34 MB slice of the filter is live in L2 at a time; the records are read with
// streaming loads (evict-first) so that they do not push that slice out.  Inside a partition the warps
// draw work items (1024 records of one region) from a counter, so nobody waits at the barrier for a
// warp that happened to own fuller regions.  Every lane carries two records per step: one 16-byte
// record load, two filter loads in flight.
#define P_SUB 1024u             // records of a region one work item covers
template <int NSTR>
__global__ void __launch_bounds__(256, 4)
probe_kernel(DevDB db, const uint64_t *__restrict__ p_words, const uint32_t *__restrict__ p_pos,
             const uint32_t *__restrict__ p_fill, uint32_t cap_cp, uint32_t n_ctas, uint32_t *__restrict__ p_ctr,
             uint64_t *__restrict__ q_words, uint32_t *__restrict__ q_slots, unsigned long long *__restrict__ q_count, uint64_t q_cap,
             uint32_t *__restrict__ hits, uint32_t *__restrict__ hitmap, unsigned long long *__restrict__ counters) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t subs = (cap_cp + P_SUB - 1) / P_SUB;            // work items per region
    const uint32_t n_items = n_ctas * subs;
    WarpQueue wq;
    wq_init(wq);
    uint32_t nh = 0;
    for (uint32_t part = 0; part < NPART; ++part) {
        for (;;) {
            uint32_t item = 0;
            if (lane == 0) item = atomicAdd(p_ctr + part, 1u);
            item = __shfl_sync(0xFFFFFFFFu, item, 0);
            if (item >= n_items) break;
            const uint32_t c = item / subs, sub = item - c * subs;
            const uint32_t fill = __ldg(p_fill + c * NPART + part);
            const uint64_t base = ((uint64_t)c * NPART + part) * cap_cp;   // even: cap_cp is even
            const uint32_t j1 = min(fill, (sub + 1) * P_SUB);
            // the records of step i+1 are requested before step i is probed (two DRAM round trips overlap)
            ulonglong2 nxt = make_ulonglong2(0, 0);
            if (sub * P_SUB + 2u * lane < j1) nxt = __ldcs(reinterpret_cast<const ulonglong2 *>(p_words + base + sub * P_SUB + 2u * lane));
            for (uint32_t j0 = sub * P_SUB; j0 < j1; j0 += 64) {
                const uint32_t j = j0 + 2u * lane;
                uint64_t w[2] = {nxt.x, nxt.y}, rc[2], hc[2];
                uint4 v[2];
                bool live[2] = {j < j1, j + 1u < j1};
                if (j + 64u < j1) nxt = __ldcs(reinterpret_cast<const ulonglong2 *>(p_words + base + j + 64u));   // j + 1 < cap_cp always: aligned 16-byte loads
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    rc[u] = revcomp_word(w[u]);
                    hc[u] = bloom_block_hash(w[u], rc[u]);
                    v[u] = make_uint4(0, 0, 0, 0);
                    if (live[u]) v[u] = __ldg(db.bloom + __umul64hi(hc[u], db.bloom_blocks));
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    // the position is only needed by survivors (~2.5 %)
                    const bool passF = live[u] && bloom_test(v[u], bloom_masks(hc[u], w[u]));
                    const bool passR = NSTR == 2 && live[u] && bloom_test(v[u], bloom_masks(hc[u], rc[u]));
                    const uint32_t bF = __ballot_sync(0xFFFFFFFFu, passF);
                    const uint32_t bR = NSTR == 2 ? __ballot_sync(0xFFFFFFFFu, passR) : 0u;
                    const uint32_t cF = __popc(bF), cnt = cF + __popc(bR);
                    if (!cnt) continue;
                    wq_reserve(wq, cnt, lane, q_slots, q_count, q_cap);
                    const uint32_t lt = (1u << lane) - 1u;
                    uint32_t pos = 0;
                    if (passF || passR) pos = __ldcs(p_pos + base + j + u);
                    if (wq.have) {
                        const uint64_t at = wq.base + wq.used;
                        if (passF) { const uint64_t i = at + __popc(bF & lt); q_words[i] = w[u]; q_slots[i] = pos * NSTR; }
                        if (passR) { const uint64_t i = at + cF + __popc(bR & lt); q_words[i] = rc[u]; q_slots[i] = pos * NSTR + 1u; }
                    } else {
                        if (passF) nh += resolve_inline(db, w[u], pos * NSTR, hits, hitmap);
                        if (passR) nh += resolve_inline(db, rc[u], pos * NSTR + 1u, hits, hitmap);
                    }
                    wq.used += cnt;
                }
            }
        }
        grid.sync();
    }
    wq_retire(wq, lane, q_slots);
    for (int o = 16; o; o >>= 1) nh += __shfl_xor_sync(0xFFFFFFFFu, nh, o);
    if (lane == 0 && nh) atomicAdd(counters + COUNTER_SLOTS + (blockIdx.x & (COUNTER_SLOTS - 1)), (unsigned long long)nh);
}

__global__ void __launch_bounds__(256, 6)
queue_lookup_kernel(DevDB db, const uint64_t *__restrict__ q_words, const uint32_t *__restrict__ q_slots,
                    const unsigned long long *__restrict__ q_count, uint64_t q_cap,
                    uint32_t *__restrict__ hits, uint32_t *__restrict__ hitmap, unsigned long long *__restrict__ counters) {
    const uint64_t n = *q_count < q_cap ? *q_count : q_cap;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t nh = 0, nsect = 0;
    // one survivor per thread (two per thread in lock-step measured slower: the pair waits for its longer chain)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t slot = q_slots[i];
        if (slot == Q_INVALID) continue;                           // padding of a retired chunk
        uint64_t ww[1] = {q_words[i]};
        uint32_t rr[1], sc;
        fast_lookup<1>(db, ww, rr, &sc);
        nsect += sc;
        if (rr[0] != HIT_MISS) { hits[slot] = rr[0]; atomicOr(hitmap + (slot >> 5), 1u << (slot & 31u)); ++nh; }
    }
    for (int o = 16; o; o >>= 1) { nh += __shfl_xor_sync(0xFFFFFFFFu, nh, o); nsect += __shfl_xor_sync(0xFFFFFFFFu, nsect, o); }
    if ((threadIdx.x & 31u) == 0) {
        const uint32_t c = blockIdx.x & (COUNTER_SLOTS - 1);
        if (nh) atomicAdd(counters + COUNTER_SLOTS + c, (unsigned long long)nh);
        if (nsect) atomicAdd(counters + 3 * COUNTER_SLOTS + c, (unsigned long long)nsect);
    }
}

// stage-level: words[] -> ix[] (utb_lookup_words)
template <bool FAST>
__global__ void lookup_words_kernel(DevDB db, const uint64_t *__restrict__ words, uint64_t n, uint32_t *__restrict__ ix) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (FAST) {
        if (db.bloom && !bloom_maybe(db, words[i])) { ix[i] = HIT_MISS; return; }
        uint64_t ww[1] = {words[i]};
        uint32_t rr[1];
        fast_lookup<1>(db, ww, rr);
        ix[i] = rr[0];
        return;
    }
    Probe q[1];
    probe_begin(db, words[i], q[0]);
    probe_run<1>(db, q);
    ix[i] = probe_end(db, q[0]);
}

// stage-level: expose every window of the packed stream (utb_pack_sequence)
__global__ void expand_windows_kernel(const uint64_t *__restrict__ pk, const uint32_t *__restrict__ bad, uint32_t n_pos,
                                      uint64_t *__restrict__ fwd, uint64_t *__restrict__ rc, uint8_t *__restrict__ valid) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n_pos) return;
    uint64_t w;
    bool ok = window_at(pk, bad, pos, w);
    fwd[pos] = ok ? w : 0;
    rc[pos] = ok ? revcomp_word(w) : 0;
    valid[pos] = ok;
}

// ---------------------------------------------------------------------------
// vote (itree.c:1028-1098)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cutoff_of(uint32_t x) {   // itree.c:1044-1046
    uint32_t c = x - x / 4u;
    c += ((x >> 1) >= c);
    return c;
}

// The aufbau walk, executed by one converged warp; every lane carries the same
// scalar state, the character scans are done 32 bytes at a time with ballots.
// T_lab/T_cnt: the distinct labels of the read in strcmp order with counts
// (shared or global memory).
__device__ void walk_warp(const DevDB &db, const uint32_t *T_lab, const uint32_t *T_cnt,
                          uint32_t uix, uint32_t n, utb_result *out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t EMPTY = 0xFFFFFFFFu;
    uint32_t cutoff = cutoff_of(n), st = 0, ed = uix, dv = EMPTY, orun = n, sl = 0, ol = 0;
    for (;;) {                                                      // itree.c:1047
        uint32_t run = T_cnt[st], td = dv;
        for (uint32_t z = st + 1; z < ed; ++z) {                    // itree.c:1050
            const char *s1 = db.blob + __ldg(db.off + T_lab[z - 1]);
            const char *s2 = db.blob + __ldg(db.off + T_lab[z]);
            if (!s1[dv + (dv == EMPTY)]) {                          // itree.c:1052
                run = T_cnt[z]; st = z;
                orun -= T_cnt[z - 1];
                cutoff = cutoff_of(orun);
                continue;
            }
            // itree.c:1060-1061: first td >= dv+1 with s1[td]==0 || s1[td]!=s2[td] || s1[td]==';'
            uint32_t t0 = dv + 1u;
            char a, b;
            for (;;) {
                char ca = s1[t0 + lane], cb = s2[t0 + lane];
                uint32_t m = __ballot_sync(0xFFFFFFFFu, ca == 0 || ca != cb || ca == ';');
                if (m) {
                    int src = __ffs(m) - 1;
                    td = t0 + (uint32_t)src;
                    a = (char)__shfl_sync(0xFFFFFFFFu, (int)ca, src);
                    b = (char)__shfl_sync(0xFFFFFFFFu, (int)cb, src);
                    break;
                }
                t0 += 32u;
            }
            if (a == b) run += T_cnt[z];                            // itree.c:1062
            else if ((!a && b == ';') ||
                     ((a == ';' || !a) && td > 0 && s1[td - 1] == '_')) {   // itree.c:1063
                run = T_cnt[z]; st = z;
                orun -= T_cnt[z - 1];
                cutoff = cutoff_of(orun);
            }
            else if (run >= cutoff) { ed = z; break; }              // itree.c:1068
            else { run = T_cnt[z]; st = z; }                        // itree.c:1069
        }
        sl = run; ol = orun;                                        // itree.c:1071
        if (run < cutoff) break;                                    // itree.c:1072
        if (st + 1 >= ed) {                                         // itree.c:1073-1080
            if (T_cnt[ed - 1] >= cutoff) dv = 0xFFFFFFFEu;
            break;
        }
        orun = run; dv = td; cutoff = cutoff_of(run);               // itree.c:1082-1085
    }
    if (lane == 0) {
        out->kind = UTB_WALK; out->label = T_lab[ed - 1]; out->cut = dv;
        out->found = n; out->uix = uix; out->sl = sl; out->ol = ol; out->_pad = 0;
    }
}

struct VoteIn {            // where a read's hit slots live
    const uint32_t *hits;
    const uint32_t *grp_off;   // batch mode: position space
    const uint32_t *seq_len;
    const uint64_t *off;       // stage-level mode: explicit ranges (overrides batch mode)
    uint32_t nstr;
    // two-phase lookup: one bit per slot that holds a label; hits[] is then only valid where the bit is set
    const uint32_t *hitmap;
};
__device__ __forceinline__ void vote_range(const VoteIn &in, uint32_t r, uint64_t &start, uint64_t &count) {
    if (in.off) { start = in.off[r]; count = in.off[r + 1] - start; return; }
    uint32_t len = __ldg(in.seq_len + r);
    uint64_t nwin = len >= 32u ? len - 31u : 0u;
    start = (uint64_t)__ldg(in.grp_off + r) * 32u * in.nstr;
    count = nwin * in.nstr;
}

#define VW_SLOTS 64u            // distinct labels a warp can hold in shared memory
#define VW_MAXHITS 16384u       // hit slots a single warp will scan
#define VW_WARPS 8

// Warp per read: hits -> (label,count) multiset in a shared-memory hash table
// (warp-aggregated with match_any), sorted by label rank, then the walk.
// Reads that are too long or have too many distinct labels are queued for
// vote_block_kernel.
struct VoteWarpSmem { uint32_t *key, *cnt, *rk, *lab, *tc; };
__device__ void vote_warp_read(const DevDB &db, const VoteIn &in, uint32_t r, utb_result *__restrict__ results,
                               uint32_t *__restrict__ gen_list, uint32_t *__restrict__ gen_count,
                               unsigned long long *__restrict__ counters, const VoteWarpSmem &sm) {
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t start, count;
    vote_range(in, r, start, count);
    utb_result *out = results + r;
    if (count > VW_MAXHITS) {
        if (lane == 0) gen_list[atomicAdd(gen_count, 1u)] = r;
        return;
    }
    uint32_t *key = sm.key, *cnt = sm.cnt;
    key[lane] = UTB_BAD32; key[lane + 32] = UTB_BAD32;
    cnt[lane] = 0; cnt[lane + 32] = 0;
    __syncwarp();
    uint32_t n = 0;
    bool overflow = false;
    if (in.hitmap) {
        // sparse: scan the read's words of the hit map (1/32 of the slot bytes), gather only the flagged slots
        const uint32_t *hm = in.hitmap + (start >> 5);             // start is a multiple of 32 in batch mode
        const uint32_t nwords = (uint32_t)((count + 31) >> 5);
        for (uint32_t wbase = 0; wbase < nwords; wbase += 32) {
            uint32_t m = wbase + lane < nwords ? __ldg(hm + wbase + lane) : 0u;
            while (__any_sync(0xFFFFFFFFu, m != 0)) {
                uint32_t h = HIT_NOWIN;
                if (m) { const uint32_t bit = __ffs(m) - 1; m &= m - 1; h = __ldg(in.hits + start + 32ull * (wbase + lane) + bit); }
                const bool ok = h < db.max_ix;
                n += __popc(__ballot_sync(0xFFFFFFFFu, ok));
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
                if (ok && (uint32_t)(__ffs(peers) - 1) == lane) {
                    uint32_t cc = __popc(peers), slot = (h * 2654435761u) >> 26;
                    for (uint32_t tries = 0;; ++tries) {
                        if (tries == VW_SLOTS) { overflow = true; break; }
                        uint32_t old = atomicCAS(&key[slot], UTB_BAD32, h);
                        if (old == UTB_BAD32 || old == h) { atomicAdd(&cnt[slot], cc); break; }
                        slot = (slot + 1) & (VW_SLOTS - 1);
                    }
                }
                __syncwarp();
            }
        }
    } else {
    // dense: 128 slots per round (one uint4 per lane); a round without a label costs one ballot
    const uint32_t mis = (uint32_t)((4u - (start & 3u)) & 3u);     // slots before the first 16-byte boundary
    for (uint64_t base = 0; base < count + 128; base += 128) {
        uint32_t hv[4];
        if (base == 0) {                                           // unaligned head: up to 3 slots, scalar
#pragma unroll
            for (int c = 0; c < 4; ++c) hv[c] = HIT_NOWIN;
            if (lane < mis && lane < count) hv[0] = __ldg(in.hits + start + lane);
        } else {
            const uint64_t i0 = mis + (base - 128) + 4ull * lane;   // aligned body
            if (i0 + 4 <= count) {
                const uint4 q4 = __ldg(reinterpret_cast<const uint4 *>(in.hits + start + i0));
                hv[0] = q4.x; hv[1] = q4.y; hv[2] = q4.z; hv[3] = q4.w;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) hv[c] = i0 + c < count ? __ldg(in.hits + start + i0 + c) : HIT_NOWIN;
            }
            if (mis + (base - 128) >= count) break;
        }
        const bool any4 = hv[0] < db.max_ix || hv[1] < db.max_ix || hv[2] < db.max_ix || hv[3] < db.max_ix;
        if (!__any_sync(0xFFFFFFFFu, any4)) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t h = hv[c];
            const bool ok = h < db.max_ix;
            const uint32_t okm = __ballot_sync(0xFFFFFFFFu, ok);
            if (!okm) continue;
            n += __popc(okm);
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
            if (ok && (uint32_t)(__ffs(peers) - 1) == lane) {      // one lane per distinct label
                uint32_t cc = __popc(peers), slot = (h * 2654435761u) >> 26;   // 6 bits
                for (uint32_t tries = 0;; ++tries) {
                    if (tries == VW_SLOTS) { overflow = true; break; }
                    uint32_t old = atomicCAS(&key[slot], UTB_BAD32, h);
                    if (old == UTB_BAD32 || old == h) { atomicAdd(&cnt[slot], cc); break; }
                    slot = (slot + 1) & (VW_SLOTS - 1);
                }
            }
            __syncwarp();
        }
    }
    }
    if (__any_sync(0xFFFFFFFFu, overflow)) {
        if (lane == 0) gen_list[atomicAdd(gen_count, 1u)] = r;
        return;
    }
    if (n == 0) {
        if (lane == 0) { out->kind = UTB_NONE; out->label = 0; out->cut = 0; out->found = 0; out->uix = 0; out->sl = 0; out->ol = 0; out->_pad = 0; }
        return;
    }
    uint32_t k0 = key[lane], k1 = key[lane + 32];
    uint32_t m0 = __ballot_sync(0xFFFFFFFFu, k0 != UTB_BAD32), m1 = __ballot_sync(0xFFFFFFFFu, k1 != UTB_BAD32);
    uint32_t uix = __popc(m0) + __popc(m1);
    if (lane == 0) atomicAdd(counters + 2 * COUNTER_SLOTS + (blockIdx.x & (COUNTER_SLOTS - 1)), 1ull);   // good finds (itree.c:1029)
    if (uix == 1) {                                                // itree.c:1031-1032, 1039-1040
        uint32_t lab = m0 ? __shfl_sync(0xFFFFFFFFu, k0, __ffs(m0) - 1) : __shfl_sync(0xFFFFFFFFu, k1, __ffs(m1) - 1);
        if (lane == 0) { out->kind = UTB_STAR; out->label = lab; out->cut = 0; out->found = n; out->uix = 1; out->sl = 0; out->ol = 0; out->_pad = 0; }
        return;
    }
    // sort the entries by label rank (strcmp order, itree.c:1041): compact the occupied slots first -- a read
    // typically holds 2-3 labels, so ranking costs uix compares per entry instead of a sweep over all 64 slots
    uint32_t *rk = sm.rk, *T_lab = sm.lab, *T_cnt = sm.tc;
    const uint32_t r0 = k0 != UTB_BAD32 ? __ldg(db.rank + k0) : UTB_BAD32;
    const uint32_t r1 = k1 != UTB_BAD32 ? __ldg(db.rank + k1) : UTB_BAD32;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t i0 = __popc(m0 & lt), i1 = __popc(m0) + __popc(m1 & lt);
    if (k0 != UTB_BAD32) rk[i0] = r0;
    if (k1 != UTB_BAD32) rk[i1] = r1;
    __syncwarp();
    uint32_t p0 = 0, p1 = 0;
    for (uint32_t j = 0; j < uix; ++j) { const uint32_t x = rk[j]; p0 += x < r0; p1 += x < r1; }
    if (k0 != UTB_BAD32) { T_lab[p0] = k0; T_cnt[p0] = cnt[lane]; }
    if (k1 != UTB_BAD32) { T_lab[p1] = k1; T_cnt[p1] = cnt[lane + 32]; }
    __syncwarp();
    walk_warp(db, T_lab, T_cnt, uix, n, out);
}
// Persistent over `list` (reads deferred by vote_thread_kernel), or over all reads when list is null.
__global__ void __launch_bounds__(VW_WARPS * 32)
vote_warp_kernel(DevDB db, VoteIn in, uint32_t n_reads, const uint32_t *__restrict__ list, const uint32_t *__restrict__ list_count,
                 utb_result *__restrict__ results, uint32_t *__restrict__ gen_list, uint32_t *__restrict__ gen_count,
                 unsigned long long *__restrict__ counters) {
    __shared__ uint32_t s_key[VW_WARPS][VW_SLOTS], s_cnt[VW_WARPS][VW_SLOTS];
    __shared__ uint32_t s_rk[VW_WARPS][VW_SLOTS], s_lab[VW_WARPS][VW_SLOTS], s_tc[VW_WARPS][VW_SLOTS];
    const uint32_t wib = threadIdx.x >> 5;
    const VoteWarpSmem sm = {s_key[wib], s_cnt[wib], s_rk[wib], s_lab[wib], s_tc[wib]};
    const uint32_t total = list ? *list_count : n_reads;
    for (uint32_t i = blockIdx.x * VW_WARPS + wib; i < total; i += gridDim.x * VW_WARPS) {
        vote_warp_read(db, in, list ? list[i] : i, results, gen_list, gen_count, counters, sm);
        __syncwarp();
    }
}

// ---- thread per read (the common case) ----------------------------------------------
// A 150-base read owns 10 words of the hit map and holds a handful of hits on 1-3 distinct labels;
// a warp per read spends ~800 issue slots on it, almost all of them with one useful lane (the
// aufbau walk compares two label strings character by character).  Here every lane walks its own
// read: hit map words -> flagged slots -> (label, count) pairs kept sorted by label rank in
// registers -> the same walk as walk_warp, scalar.  Reads with more than VT_K distinct labels or
// more than VT_MAXWORDS hit-map words are deferred to vote_warp_kernel through `list`.
#define VT_K 6u
#define VT_THREADS 128
#define VT_MAXWORDS 64u
__device__ __forceinline__ void walk_thread(const DevDB &db, const uint32_t *T_lab, const uint32_t *T_cnt,   // [k * VT_THREADS]
                                            uint32_t uix, uint32_t n, utb_result *out) {
#define TL(z) T_lab[(z) * VT_THREADS]
#define TC(z) T_cnt[(z) * VT_THREADS]
    const uint32_t EMPTY = 0xFFFFFFFFu;
    uint32_t cutoff = cutoff_of(n), st = 0, ed = uix, dv = EMPTY, orun = n, sl = 0, ol = 0;
    for (;;) {                                                      // itree.c:1047
        uint32_t run = TC(st), td = dv;
        for (uint32_t z = st + 1; z < ed; ++z) {                    // itree.c:1050
            const char *s1 = db.blob + __ldg(db.off + TL(z - 1));
            const char *s2 = db.blob + __ldg(db.off + TL(z));
            if (!s1[dv + (dv == EMPTY)]) {                          // itree.c:1052
                run = TC(z); st = z;
                orun -= TC(z - 1);
                cutoff = cutoff_of(orun);
                continue;
            }
            // itree.c:1060-1061: first td >= dv+1 with s1[td]==0 || s1[td]!=s2[td] || s1[td]==';'
            uint32_t t = dv + 1u;
            char a, b;
            for (;; ++t) { a = s1[t]; b = s2[t]; if (a == 0 || a != b || a == ';') break; }
            td = t;
            if (a == b) run += TC(z);                               // itree.c:1062
            else if ((!a && b == ';') ||
                     ((a == ';' || !a) && td > 0 && s1[td - 1] == '_')) {   // itree.c:1063
                run = TC(z); st = z;
                orun -= TC(z - 1);
                cutoff = cutoff_of(orun);
            }
            else if (run >= cutoff) { ed = z; break; }              // itree.c:1068
            else { run = TC(z); st = z; }                           // itree.c:1069
        }
        sl = run; ol = orun;                                        // itree.c:1071
        if (run < cutoff) break;                                    // itree.c:1072
        if (st + 1 >= ed) {                                         // itree.c:1073-1080
            if (TC(ed - 1) >= cutoff) dv = 0xFFFFFFFEu;
            break;
        }
        orun = run; dv = td; cutoff = cutoff_of(run);               // itree.c:1082-1085
    }
    out->kind = UTB_WALK; out->label = TL(ed - 1); out->cut = dv;
    out->found = n; out->uix = uix; out->sl = sl; out->ol = ol; out->_pad = 0;
#undef TL
#undef TC
}
__global__ void __launch_bounds__(VT_THREADS)
vote_thread_kernel(DevDB db, VoteIn in, uint32_t n_reads, utb_result *__restrict__ results,
                   uint32_t *__restrict__ list, uint32_t *__restrict__ list_count, unsigned long long *__restrict__ counters) {
    __shared__ uint32_t s_lab[VT_K * VT_THREADS], s_cnt[VT_K * VT_THREADS];
    __shared__ uint32_t s_good;
    const uint32_t tid = threadIdx.x, r = blockIdx.x * VT_THREADS + tid;
    if (tid == 0) s_good = 0;
    __syncthreads();
    if (r < n_reads) {
        uint64_t start, count;
        vote_range(in, r, start, count);
        const uint32_t nwords = (uint32_t)((count + 31) >> 5);
        bool defer = nwords > VT_MAXWORDS;
        uint32_t lab[VT_K], cnt[VT_K], n = 0, uix = 0;
#pragma unroll
        for (uint32_t k = 0; k < VT_K; ++k) { lab[k] = UTB_BAD32; cnt[k] = 0; }
        if (!defer) {
            const uint32_t *hm = in.hitmap + (start >> 5);         // start is a multiple of 32 in batch mode
            for (uint32_t w = 0; w < nwords && !defer; ++w) {
                uint32_t m = __ldg(hm + w);
                if (w == nwords - 1 && (count & 31u)) m &= (1u << (count & 31u)) - 1u;
                while (m) {
                    const uint32_t bit = __ffs(m) - 1; m &= m - 1;
                    const uint32_t h = __ldg(in.hits + start + 32ull * w + bit);
                    if (h >= db.max_ix) continue;
                    ++n;
                    bool seen = false;
#pragma unroll
                    for (uint32_t k = 0; k < VT_K; ++k) if (lab[k] == h) { ++cnt[k]; seen = true; }
                    if (!seen) {
                        if (uix == VT_K) { defer = true; break; }
#pragma unroll
                        for (uint32_t k = 0; k < VT_K; ++k) if (k == uix) { lab[k] = h; cnt[k] = 1; }
                        ++uix;
                    }
                }
            }
        }
        utb_result *out = results + r;
        if (defer) list[atomicAdd(list_count, 1u)] = r;
        else if (n == 0) { out->kind = UTB_NONE; out->label = 0; out->cut = 0; out->found = 0; out->uix = 0; out->sl = 0; out->ol = 0; out->_pad = 0; }
        else {
            atomicAdd(&s_good, 1u);                                 // good finds (itree.c:1029)
            if (uix == 1) {                                         // itree.c:1031-1032, 1039-1040
                out->kind = UTB_STAR; out->label = lab[0]; out->cut = 0; out->found = n; out->uix = 1; out->sl = 0; out->ol = 0; out->_pad = 0;
            } else {
                // sort by label rank (strcmp order, itree.c:1041): insertion sort over at most VT_K entries
                uint32_t rk[VT_K];
#pragma unroll
                for (uint32_t k = 0; k < VT_K; ++k) rk[k] = k < uix ? __ldg(db.rank + lab[k]) : UTB_BAD32;
#pragma unroll
                for (uint32_t i = 1; i < VT_K; ++i)
#pragma unroll
                    for (uint32_t j = i; j > 0; --j)
                        if (rk[j] < rk[j - 1]) {
                            uint32_t t = rk[j]; rk[j] = rk[j - 1]; rk[j - 1] = t;
                            t = lab[j]; lab[j] = lab[j - 1]; lab[j - 1] = t;
                            t = cnt[j]; cnt[j] = cnt[j - 1]; cnt[j - 1] = t;
                        }
#pragma unroll
                for (uint32_t k = 0; k < VT_K; ++k) { s_lab[k * VT_THREADS + tid] = lab[k]; s_cnt[k * VT_THREADS + tid] = cnt[k]; }
                walk_thread(db, s_lab + tid, s_cnt + tid, uix, n, out);
            }
        }
    }
    __syncthreads();
    if (tid == 0 && s_good) atomicAdd(counters + 2 * COUNTER_SLOTS + (blockIdx.x & (COUNTER_SLOTS - 1)), (unsigned long long)s_good);
}

// Block per read for long queries / label-rich reads: global-memory histogram
// over label ids (block-private scratch, kept zeroed), compaction in rank
// order, then the same walk.  Persistent over the queue filled by
// vote_warp_kernel.
#define VB_THREADS 256
__global__ void __launch_bounds__(VB_THREADS)
vote_block_kernel(DevDB db, VoteIn in, utb_result *__restrict__ results,
                  const uint32_t *__restrict__ gen_list, const uint32_t *__restrict__ gen_count,
                  uint32_t *__restrict__ hist_all, uint32_t *__restrict__ tlab_all, uint32_t *__restrict__ tcnt_all,
                  unsigned long long *__restrict__ counters) {
    __shared__ uint32_t s_warp[VB_THREADS / 32];
    __shared__ uint32_t s_base, s_n;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
    uint32_t *hist = hist_all + (size_t)blockIdx.x * db.max_ix;
    uint32_t *T_lab = tlab_all + (size_t)blockIdx.x * db.max_ix;
    uint32_t *T_cnt = tcnt_all + (size_t)blockIdx.x * db.max_ix;
    const uint32_t total = *gen_count;
    for (uint32_t qi = blockIdx.x; qi < total; qi += gridDim.x) {
        const uint32_t r = gen_list[qi];
        uint64_t start, count;
        vote_range(in, r, start, count);
        // 1. histogram (warp-aggregated global atomics) + foundUniq
        uint32_t n_local = 0;
        if (in.hitmap) {
            const uint32_t *hm = in.hitmap + (start >> 5);
            const uint64_t nwords = (count + 31) >> 5;
            for (uint64_t wi = tid; wi < nwords; wi += VB_THREADS) {
                uint32_t m = __ldg(hm + wi);
                while (m) {
                    const uint32_t bit = __ffs(m) - 1; m &= m - 1;
                    const uint32_t h = __ldg(in.hits + start + 32ull * wi + bit);
                    if (h < db.max_ix) { atomicAdd(&hist[h], 1u); ++n_local; }
                }
            }
        } else
        for (uint64_t base = 0; base < count; base += VB_THREADS) {
            uint64_t i = base + tid;
            uint32_t h = i < count ? __ldg(in.hits + start + i) : HIT_NOWIN;
            bool ok = h < db.max_ix;
            uint32_t peers = __match_any_sync(0xFFFFFFFFu, h);
            if (ok && (uint32_t)(__ffs(peers) - 1) == lane) atomicAdd(&hist[h], (uint32_t)__popc(peers));
            n_local += ok;
        }
        for (int o = 16; o; o >>= 1) n_local += __shfl_xor_sync(0xFFFFFFFFu, n_local, o);
        if (tid == 0) { s_n = 0; s_base = 0; }
        __syncthreads();
        if (lane == 0 && n_local) atomicAdd(&s_n, n_local);
        __threadfence();
        __syncthreads();
        const uint32_t n = s_n;
        // 2. compaction in rank order (itree.c:1036-1041 produce exactly this list)
        for (uint32_t r0 = 0; r0 < db.max_ix; r0 += VB_THREADS) {
            uint32_t rr = r0 + tid;
            uint32_t lab = rr < db.max_ix ? __ldg(db.by_rank + rr) : 0;
            uint32_t c = rr < db.max_ix ? __ldcg(hist + lab) : 0;    // L2 view: the counts were made by atomics
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, c != 0);
            if (lane == 0) s_warp[wid] = __popc(bal);
            __syncthreads();
            uint32_t wbase = s_base;
            for (uint32_t w = 0; w < wid; ++w) wbase += s_warp[w];
            if (c) {
                uint32_t p = wbase + __popc(bal & ((1u << lane) - 1u));
                T_lab[p] = lab; T_cnt[p] = c;
                hist[lab] = 0;                                     // leave the scratch clean
            }
            __syncthreads();
            if (tid == 0) { uint32_t t = 0; for (uint32_t w = 0; w < VB_THREADS / 32; ++w) t += s_warp[w]; s_base += t; }
            __syncthreads();
        }
        __threadfence();
        __syncthreads();
        const uint32_t uix = s_base;
        utb_result *out = results + r;
        if (wid == 0) {
            if (n == 0) {
                if (lane == 0) { out->kind = UTB_NONE; out->label = 0; out->cut = 0; out->found = 0; out->uix = 0; out->sl = 0; out->ol = 0; out->_pad = 0; }
            } else {
                if (lane == 0) atomicAdd(counters + 2 * COUNTER_SLOTS + (blockIdx.x & (COUNTER_SLOTS - 1)), 1ull);
                if (uix == 1) {
                    if (lane == 0) { out->kind = UTB_STAR; out->label = T_lab[0]; out->cut = 0; out->found = n; out->uix = 1; out->sl = 0; out->ol = 0; out->_pad = 0; }
                } else walk_warp(db, T_lab, T_cnt, uix, n, out);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// output text on the device (the fprintf lines of itree.c:1032, 1040, 1096)
// ---------------------------------------------------------------------------
// The host formatter competes with the framer for a handful of cores (16 for
// 8 GPUs on the test box), so the lines are produced here: per-read length,
// exclusive scan, then one warp per read copies name, label prefix and the
// numeric tail.  The host only moves the finished text to the sink.
__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
    return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u :
           v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}
__device__ __forceinline__ uint32_t tax_len_of(const DevDB &db, const utb_result &v) {
    uint32_t ll = __ldg(db.off + v.label + 1) - __ldg(db.off + v.label) - 1u;
    if (v.kind == UTB_WALK) {
        if (v.cut == UTB_CUT_EMPTY) ll = 0;                        // dv == -1 (itree.c:1087)
        else if (v.cut != UTB_CUT_FULL && v.cut < ll) ll = v.cut;  // first dv bytes (itree.c:1088)
    }
    return ll;
}
__global__ void __launch_bounds__(256)
fmt_len_kernel(DevDB db, const utb_result *__restrict__ res, const uint32_t *__restrict__ name_len, uint32_t n_reads,
               uint32_t *__restrict__ line_len) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const utb_result v = res[r];
    uint32_t n = 0;
    if (v.kind != UTB_NONE) {
        n = __ldg(name_len + r) + 1u + tax_len_of(db, v) + 1u + dec_digits(v.found) + 1u + dec_digits(v.uix) + 1u;
        n += v.kind == UTB_STAR ? 1u : dec_digits(v.sl) + 1u + dec_digits(v.ol);
        n += 1u;
    }
    line_len[r] = n;
}
// exclusive scan of u32 (n up to 2^31): block sums, scan of the sums by one block, local scan + base
#define SCAN_TILE 2048u
__global__ void __launch_bounds__(256)
scan_sums_kernel(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ sums) {
    __shared__ uint32_t sh[8];
    const uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < SCAN_TILE; i += 256) if (base + i < n) acc += in[base + i];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; sums[blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024)
scan_top_kernel(uint32_t *__restrict__ sums, uint32_t nb, uint32_t *__restrict__ total) {
    __shared__ uint32_t sh[1024];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nb; base += 1024) {
        uint32_t i = base + threadIdx.x, v = i < nb ? sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t o = 1; o < 1024; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) sums[i] = carry + sh[threadIdx.x] - v;           // exclusive
        carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256)
scan_apply_kernel(const uint32_t *__restrict__ in, uint32_t n, const uint32_t *__restrict__ sums, uint32_t *__restrict__ out) {
    __shared__ uint32_t sh[8];
    __shared__ uint32_t run;
    const uint32_t base = blockIdx.x * SCAN_TILE;
    if (threadIdx.x == 0) run = sums[blockIdx.x];
    __syncthreads();
    for (uint32_t c = 0; c < SCAN_TILE; c += 256) {
        const uint32_t i = base + c + threadIdx.x;
        const uint32_t v = i < n ? in[i] : 0;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o); if ((threadIdx.x & 31) >= (uint32_t)o) x += t; }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t wbase = run;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += sh[w];
        if (i < n) out[i] = wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 255) run = wbase + x;
        __syncthreads();
    }
}
__device__ __forceinline__ uint32_t put_dec(char *p, uint32_t v) {
    const uint32_t n = dec_digits(v);
    for (uint32_t i = n; i; --i) { p[i - 1] = (char)('0' + v % 10u); v /= 10u; }
    return n;
}
__global__ void __launch_bounds__(256)
slots_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, uint32_t *__restrict__ slots) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) slots[r] = (seq_len[r] + 1u + 31u) / 32u;           // utb_read_slots
}
// ---- device-side framing ---------------------------------------------
// The host only counts the newlines of a chunk (which fixes the number of complete
// records and where the last one ends, itree.c:869-871: lines are read strictly in
// pairs); everything per read -- where its header and sequence lines start, the name
// (itree.c:879-882), the CR/LF trimming (itree.c:889-890) and the format checks
// (itree.c:880, 886) -- happens here, on the bytes that were copied to the device
// anyway.  Three kernels: newline count per 16 KB block, exclusive scan of the block
// counts (scan_top_kernel), newline positions + record parse.
#define FR_BPT 64u              // bytes per thread
#define FR_BLOCK (256u * FR_BPT)
#define FR_ERR_NONE 0xFFFFFFFFu
enum { FRE_NOHEADER = 1, FRE_SEQ_GT = 2, FRE_TOOLONG = 4 };   // codes as in pipeline.c (FE_*)
// bit j of the result: byte j of the 64 bytes at p is '\n' (bytes at or beyond n_bytes never match);
// *nul (optional): whether one of those bytes is 0
__device__ __forceinline__ uint64_t nl_mask64(const uint8_t *__restrict__ raw, uint64_t at, uint64_t n_bytes, bool *nul = nullptr) {
    uint64_t m = 0, z = 0;
    if (nul) *nul = false;
    if (at >= n_bytes) return 0;
    const uint4 *p = reinterpret_cast<const uint4 *>(raw + at);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldg(p + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t e = __vcmpeq4(w[k], 0x0A0A0A0Au) & 0x01010101u;
            m |= (uint64_t)((e * 0x01020408u) >> 24) << (16 * q + 4 * k);
            if (nul) { const uint32_t e0 = __vcmpeq4(w[k], 0u) & 0x01010101u; z |= (uint64_t)((e0 * 0x01020408u) >> 24) << (16 * q + 4 * k); }
        }
    }
    const uint64_t left = n_bytes - at;
    if (left < 64) { m &= (1ull << left) - 1ull; z &= (1ull << left) - 1ull; }
    if (nul) *nul = z != 0;
    return m;
}
__global__ void __launch_bounds__(256)
nl_count_kernel(const uint8_t *__restrict__ raw, uint64_t n_bytes, uint32_t *__restrict__ blk_cnt, uint32_t *__restrict__ nul_flag) {
    __shared__ uint32_t sh[8];
    const uint64_t at = ((uint64_t)blockIdx.x * 256u + threadIdx.x) * FR_BPT;
    bool nul;
    uint32_t c = __popcll(nl_mask64(raw, at, n_bytes, &nul));
    if (__any_sync(0xFFFFFFFFu, nul) && (threadIdx.x & 31) == 0) atomicOr(nul_flag, 1u);
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += sh[w]; blk_cnt[blockIdx.x] = t; }
}
// blk_off: exclusive scan of blk_cnt.  nl[i] = position of newline i, for i < limit.
__global__ void __launch_bounds__(256)
nl_index_kernel(const uint8_t *__restrict__ raw, uint64_t n_bytes, const uint32_t *__restrict__ blk_off,
                uint32_t limit, uint32_t *__restrict__ nl) {
    __shared__ uint32_t sh[8];
    const uint64_t at = ((uint64_t)blockIdx.x * 256u + threadIdx.x) * FR_BPT;
    uint64_t m = nl_mask64(raw, at, n_bytes);
    const uint32_t c = __popcll(m), lane = threadIdx.x & 31u;
    uint32_t x = c;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= (uint32_t)o) x += t; }
    if (lane == 31) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t i = blk_off[blockIdx.x] + x - c;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) i += sh[w];
    while (m && i < limit) {
        nl[i++] = (uint32_t)(at + (uint32_t)(__ffsll((long long)m) - 1));
        m &= m - 1;
    }
}
// One thread per record r: header line = (nl[2r-1], nl[2r]], sequence line = (nl[2r], nl[2r+1]].
// The first malformed record (smallest r) is reported in *err as r << 3 | code; the host then
// re-frames from this batch on with its own exact reader, which reproduces the reference's message
// and partial output.
__global__ void __launch_bounds__(256)
frame_parse_kernel(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ nl, uint32_t n_reads,
                   uint64_t *__restrict__ seq_off, uint32_t *__restrict__ seq_len,
                   uint32_t *__restrict__ name_off, uint32_t *__restrict__ name_len,
                   uint32_t *__restrict__ slots, uint32_t *__restrict__ err) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t hs = r ? nl[2 * r - 1] + 1u : 0u, hn = nl[2 * r], ss = hn + 1u, sn = nl[2 * r + 1];
    uint32_t code = 0;
    if (ss - hs >= UTB_LINELEN || sn + 1u - ss >= UTB_LINELEN) code = FRE_TOOLONG;
    else if (raw[hs] != '>') code = FRE_NOHEADER;                  // itree.c:880
    else if (raw[ss] == '>') code = FRE_SEQ_GT;                    // itree.c:886
    if (code) atomicMin(err, (r << 3) | code);
    uint32_t length = sn - ss;
    if (length && raw[ss + length - 1] == '\r') --length;          // itree.c:890
    uint32_t e = hs + 1u;                                          // name: up to the first ' ' or the newline (itree.c:881)
    while (e < hn && raw[e] != ' ') ++e;
    seq_off[r] = ss; seq_len[r] = length;
    name_off[r] = hs + 1u; name_len[r] = e - hs - 1u;
    slots[r] = (length + 1u + 31u) / 32u;                          // utb_read_slots
}

#define FW_WARPS 8
__global__ void __launch_bounds__(FW_WARPS * 32)
fmt_write_kernel(DevDB db, const utb_result *__restrict__ res, const uint8_t *__restrict__ raw,
                 const uint32_t *__restrict__ name_off, const uint32_t *__restrict__ name_len,
                 const uint32_t *__restrict__ line_off, uint32_t n_reads, char *__restrict__ text) {
    __shared__ char s_tail[FW_WARPS][48];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * FW_WARPS + wib;
    if (r >= n_reads) return;
    const utb_result v = res[r];
    if (v.kind == UTB_NONE) return;
    char *out = text + line_off[r];
    const uint32_t nl = __ldg(name_len + r), tl = tax_len_of(db, v);
    const uint8_t *nm = raw + __ldg(name_off + r);
    const char *lab = db.blob + __ldg(db.off + v.label);
    uint32_t tn = 0;
    if (lane == 0) {                                               // "\t<found>\t<uix>\t*\n" or "...\t<sl>;<ol>\n"
        char *t = s_tail[wib];
        t[tn++] = '\t'; tn += put_dec(t + tn, v.found);
        t[tn++] = '\t'; tn += put_dec(t + tn, v.uix);
        t[tn++] = '\t';
        if (v.kind == UTB_STAR) t[tn++] = '*';
        else { tn += put_dec(t + tn, v.sl); t[tn++] = ';'; tn += put_dec(t + tn, v.ol); }
        t[tn++] = '\n';
    }
    tn = __shfl_sync(0xFFFFFFFFu, tn, 0);
    __syncwarp();
    for (uint32_t i = lane; i < nl; i += 32) out[i] = (char)nm[i];
    if (lane == 0) out[nl] = '\t';
    for (uint32_t i = lane; i < tl; i += 32) out[nl + 1 + i] = lab[i];
    for (uint32_t i = lane; i < tn; i += 32) out[nl + 1 + tl + i] = s_tail[wib][i];
}

// ---------------------------------------------------------------------------
// random-sector gather microbenchmark (roofline denominator)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rand32_kernel(const uint8_t *__restrict__ buf, uint64_t n_sectors, uint32_t per_thread, uint64_t seed,
              unsigned long long *__restrict__ sink) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t x = (t + 1) * 0x9E3779B97F4A7C15ull + seed;
    uint64_t acc = 0;
    for (uint32_t k = 0; k < per_thread; k += 8) {
        uint64_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;              // xorshift64*
            uint64_t s = __umul64hi(x * 0x2545F4914F6CDD1Dull, n_sectors);   // uniform in [0, n_sectors)
            v[j] = __ldg(reinterpret_cast<const uint64_t *>(buf + s * 32));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += v[j];
    }
    if (acc == 0x123456789ull) atomicAdd(sink, 1ull);              // keep the loads alive
}

// ---------------------------------------------------------------------------
// C ABI: database residency
// ---------------------------------------------------------------------------
extern "C" int utb_device_count(int *n) {
    if (!n) { utb_set_error("utb_device_count: null argument"); return UTB_ERR_ARG; }
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess || *n <= 0) {
        *n = 0;
        utb_set_error("no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
        return UTB_ERR_CUDA;
    }
    return UTB_OK;
}

static int check_device(int device) {
    int n = 0;
    int rc = utb_device_count(&n);
    if (rc) return rc;
    if (device < 0 || device >= n) { utb_set_error("device %d out of range (have %d)", device, n); return UTB_ERR_ARG; }
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) {
        utb_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
        return UTB_ERR_CUDA;
    }
    return UTB_OK;
}

// host (pageable / mmap'ed) -> device through two pinned bounce buffers so the
// page-cache copy of chunk i+1 overlaps the PCIe copy of chunk i
static int upload_streamed(void *dst, const void *src, size_t n, cudaStream_t st) {
    const size_t CH = (size_t)64 << 20;
    void *pin[2] = {nullptr, nullptr};
    cudaEvent_t ev[2];
    CK(cudaMallocHost(&pin[0], CH));
    CK(cudaMallocHost(&pin[1], CH));
    CK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int k = 0;
    for (size_t o = 0; o < n; o += CH, k ^= 1) {
        size_t c = n - o < CH ? n - o : CH;
        CK(cudaEventSynchronize(ev[k]));
        memcpy(pin[k], (const char *)src + o, c);
        CK(cudaMemcpyAsync((char *)dst + o, pin[k], c, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ev[k], st));
    }
    CK(cudaStreamSynchronize(st));
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    cudaFreeHost(pin[0]); cudaFreeHost(pin[1]);
    return UTB_OK;
}

static int db_upload_impl(const utb_ctr *ctr, int device, utb_db *db);
extern "C" int utb_db_upload(const utb_ctr *ctr, int device, utb_db **out) {
    if (!ctr || !out) { utb_set_error("utb_db_upload: null argument"); return UTB_ERR_ARG; }
    *out = nullptr;
    int rc = check_device(device);
    if (rc) return rc;
    CK(cudaSetDevice(device));
    utb_db *db = (utb_db *)calloc(1, sizeof(utb_db));
    if (!db) { utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    db->device = device;
    rc = db_upload_impl(ctr, device, db);
    if (rc) { utb_db_free(db); return rc; }                        // device buffers allocated so far are released
    *out = db;
    return UTB_OK;
}
static int db_upload_impl(const utb_ctr *ctr, int device, utb_db *db) {
    int rc;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    size_t nb_binix = (size_t)UTB_NUMBINS * ctr->binix_bytes;
    size_t nb_recs = (size_t)(ctr->num_nodes * ctr->sz);
    size_t nl = ctr->max_ix;
    CK(cudaMalloc(&db->binix, nb_binix + 64));
    CK(cudaMalloc(&db->recs, nb_recs + 64));                       // +slack: itree.c:766
    CK(cudaMalloc(&db->blob, ctr->blob_len + 64));
    CK(cudaMalloc(&db->off, (nl + 1) * 4));
    CK(cudaMalloc(&db->rank, (nl + 1) * 4));
    CK(cudaMalloc(&db->by_rank, (nl + 1) * 4));
    CK(cudaMemsetAsync((char *)db->binix + nb_binix, 0, 64, st));
    CK(cudaMemsetAsync((char *)db->recs + nb_recs, 0, 64, st));
    CK(cudaMemsetAsync(db->blob, 0, ctr->blob_len + 64, st));
    rc = upload_streamed(db->binix, ctr->binix_raw, nb_binix, st); if (rc) return rc;
    rc = upload_streamed(db->recs, ctr->recs, nb_recs, st); if (rc) return rc;
    CK(cudaMemcpyAsync(db->blob, ctr->blob, ctr->blob_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db->off, ctr->off, (nl + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db->rank, ctr->rank, nl * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(db->by_rank, ctr->by_rank, nl * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamDestroy(st));
    db->d.binix32 = ctr->binix_bytes == 4 ? (const uint32_t *)db->binix : nullptr;
    db->d.binix64 = ctr->binix_bytes == 8 ? (const uint64_t *)db->binix : nullptr;
    db->d.recs = (const uint8_t *)db->recs;
    db->d.num_nodes = ctr->num_nodes;
    db->d.sz = ctr->sz; db->d.ix_bytes = ctr->ix_bytes; db->d.max_ix = ctr->max_ix;
    db->d.blob = (const char *)db->blob;
    db->d.off = (const uint32_t *)db->off;
    db->d.rank = (const uint32_t *)db->rank;
    db->d.by_rank = (const uint32_t *)db->by_rank;
    db->hbm_bytes = nb_binix + nb_recs + ctr->blob_len + 3 * (nl + 1) * 4;
    for (size_t i = 0; i < nl; ++i) { size_t l = ctr->off[i + 1] - ctr->off[i]; if (l > db->max_label) db->max_label = l; }
    // Is the CTR regular (what utree-compress emits: every bucket strictly
    // sorted, at most the first-bin quirk)?  Then the interpolation search is
    // exact; otherwise keep the reference's probe sequence.
    {
        db->d.quirk_bin = 0xFFFFFFFFu;
        unsigned long long *d_out, h_out[4] = {0, 0, ~0ull, ~0ull};
        CK(cudaMalloc(&d_out, 32));
        CK(cudaMemcpy(d_out, h_out, 32, cudaMemcpyHostToDevice));
        verify_kernel<<<(UTB_NUMBINS - 1 + 255) / 256, 256>>>(db->d, d_out);
        CK(cudaGetLastError());
        CK(cudaMemcpy(h_out, d_out, 32, cudaMemcpyDeviceToHost));
        cudaFree(d_out);
        db->regular = h_out[0] == 0 && (h_out[1] == 0 || (h_out[1] == 1 && h_out[2] == h_out[3]));
        if (db->regular && h_out[1] == 1) db->d.quirk_bin = (uint32_t)h_out[2];
        const char *lk = getenv("UTB_LOOKUP");                      // "exact" forces the reference probe sequence
        db->use_interp = db->regular && !(lk && !strcmp(lk, "exact"));
    }
    if (db->use_interp) {
        // one-time re-layout (slack after the last record: a window load may overrun into it)
        const size_t n = (size_t)ctr->num_nodes;
        size_t lay_bytes;
        if (ctr->ix_bytes == 2) {                                  // key/aux interleaved per 128-byte line
            const size_t blocks = (n + 15) / 16 + 2;
            lay_bytes = blocks * 128;
            CK(cudaMalloc(&db->keys, lay_bytes));
            CK(cudaMemset(db->keys, 0xFF, lay_bytes));
            relayout_kernel<<<(unsigned)((n + 255) / 256), 256>>>(db->d, (uint32_t *)db->keys, nullptr, nullptr);
        } else {
            lay_bytes = (n + 32) * 12;
            CK(cudaMalloc(&db->keys, (n + 32) * 4));
            CK(cudaMalloc(&db->aux, (n + 32) * 8));
            CK(cudaMemset((char *)db->keys + n * 4, 0xFF, 32 * 4));
            relayout_kernel<<<(unsigned)((n + 255) / 256), 256>>>(db->d, nullptr, (uint32_t *)db->keys, (uint64_t *)db->aux);
        }
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        db->d.kv = ctr->ix_bytes == 2 ? (const uint32_t *)db->keys : nullptr;
        db->d.keys = ctr->ix_bytes == 2 ? nullptr : (const uint32_t *)db->keys;
        db->d.aux64 = ctr->ix_bytes == 2 ? nullptr : (const uint64_t *)db->aux;
        CK(cudaFree(db->recs));                                    // the byte-packed blob is no longer needed
        db->recs = nullptr; db->d.recs = nullptr;
        db->hbm_bytes = nb_binix + lay_bytes + ctr->blob_len + 3 * (nl + 1) * 4;
        const char *bm = getenv("UTB_BLOOM");                      // 0 off, 1 always, default auto
        db->bloom_mode = bm ? (atoi(bm) == 0 ? 0 : atoi(bm) == 1 ? 1 : 2) : 2;
        if (db->bloom_mode) {
            // 16 bits per record: 8 records per 128-bit block -> ~0.1 % false positives
            uint64_t blocks = n / 8 + 1024;
            CK(cudaMalloc(&db->bloom, blocks * 16));
            CK(cudaMemset(db->bloom, 0, blocks * 16));
            db->d.bloom_blocks = blocks;
            bloom_build_kernel<<<(UTB_NUMBINS - 1 + 255) / 256, 256>>>(db->d, (uint32_t *)db->bloom);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            db->d.bloom = (const uint4 *)db->bloom;
            db->hbm_bytes += blocks * 16;
        }
    }
    // Hot prefix table pinned in L2: reserve persisting lines for the index so
    // the streaming record traffic does not evict it (applied per stream in
    // utb_batch_create).
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    const char *env = getenv("UTB_L2_PERSIST");
    if ((!env || atoi(env) != 0) && p.persistingL2CacheMaxSize > 0 && p.accessPolicyMaxWindowSize > 0) {
        size_t want = nb_binix < (size_t)p.persistingL2CacheMaxSize ? nb_binix : (size_t)p.persistingL2CacheMaxSize;
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            db->l2_window = 1;
            db->l2_window_bytes = nb_binix < (size_t)p.accessPolicyMaxWindowSize ? nb_binix : (size_t)p.accessPolicyMaxWindowSize;
            // a window larger than the set-aside would thrash it: persist only the fraction that fits
            db->l2_hit_ratio = want >= db->l2_window_bytes ? 1.0f : (float)((double)want / (double)db->l2_window_bytes);
            const char *hr = getenv("UTB_L2_HITRATIO");
            if (hr) db->l2_hit_ratio = (float)atof(hr);
        } else cudaGetLastError();
    }
    const char *fg = getenv("UTB_L2_FETCH");                        // DRAM->L2 fetch granularity hint: 32 / 64 / 128
    if (fg && atoi(fg) > 0 && cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg)) != cudaSuccess) cudaGetLastError();
    if (getenv("UTB_STATS")) {
        size_t fgv = 0, pl = 0;
        cudaDeviceGetLimit(&fgv, cudaLimitMaxL2FetchGranularity);
        cudaDeviceGetLimit(&pl, cudaLimitPersistingL2CacheSize);
        fprintf(stderr, "utree-b200: device %d %s: L2 %d MiB, persisting max %d MiB (set %zu MiB), window max %d MiB, "
                "fetch granularity %zu B, binix window %zu MiB hitRatio %.2f\n", device, p.name, p.l2CacheSize >> 20,
                p.persistingL2CacheMaxSize >> 20, pl >> 20, p.accessPolicyMaxWindowSize >> 20, fgv,
                db->l2_window_bytes >> 20, db->l2_hit_ratio);
    }
    return UTB_OK;
}

extern "C" void utb_db_free(utb_db *db) {
    if (!db) return;
    cudaSetDevice(db->device);
    cudaFree(db->binix); cudaFree(db->recs); cudaFree(db->blob); cudaFree(db->keys); cudaFree(db->aux); cudaFree(db->bloom);
    cudaFree(db->off); cudaFree(db->rank); cudaFree(db->by_rank);
    free(db);
}
extern "C" uint64_t utb_db_hbm_bytes(const utb_db *db) { return db ? db->hbm_bytes : 0; }
extern "C" int utb_db_lookup_mode(const utb_db *db) { return db ? db->use_interp : 0; }

// ---------------------------------------------------------------------------
// C ABI: batches
// ---------------------------------------------------------------------------
#define VB_BLOCKS 148

struct utb_batch {
    utb_db *db;
    cudaStream_t st;
    cudaEvent_t done, ev[7];
    size_t max_bytes, max_reads;
    uint64_t max_groups;
    // pinned host
    char *h_bytes; uint64_t *h_seq_off; uint32_t *h_seq_len; uint32_t *h_grp_off;
    utb_result *h_results; unsigned long long *h_counters;
    uint32_t *h_name_off, *h_name_len; char *h_text; uint32_t *h_text_len; size_t text_cap;   // device-side formatting
    // device
    uint8_t *d_raw; uint64_t *d_seq_off; uint32_t *d_seq_len; uint32_t *d_grp_off;
    uint64_t *d_pk; uint32_t *d_bad; uint32_t *d_hits;
    utb_result *d_results; uint32_t *d_gen_list; uint32_t *d_gen_count; uint32_t *d_warp_list; uint32_t *d_warp_count;
    uint32_t *d_name_off, *d_name_len, *d_line_len, *d_line_off, *d_scan_sums, *d_text_len; char *d_text;
    int want_text; cudaEvent_t text_len_ready; size_t text_prefetched;
    unsigned long long *d_counters;   // [4][COUNTER_SLOTS]: lookups, hits, good finds, exact-path sectors (summed on the host)
    uint32_t *d_hist, *d_tlab, *d_tcnt;
    uint64_t *d_qwords; uint32_t *d_qslots; unsigned long long *d_qcount; uint64_t q_cap;   // filter survivors
    uint32_t *d_hitmap;
    uint64_t *d_pwords; uint32_t *d_ppos; uint32_t *d_pfill; uint32_t *d_pctr; uint64_t p_total, part_min_pos;   // partitioned filter pass
    // last submit
    size_t n_reads; uint32_t n_groups; int do_rc; int in_flight; int used_bloom; int used_partition;
    uint64_t launches;
    // device-side framing
    uint32_t *d_nl, *d_blk, *d_frame_err, *d_ngroups, *h_frame_err; const uint32_t *ngroups_dev; int framed_on_device;
    uint32_t *d_frame_info, *h_frame_info; cudaEvent_t count_ready; size_t raw_bytes;   // [0] newlines of the chunk, [1] NUL seen
};

extern "C" uint64_t utb_read_slots(uint32_t len) { return ((uint64_t)len + 1 + 31) / 32; }
extern "C" uint64_t utb_batch_max_slots(const utb_batch *b) { return b->max_groups; }
extern "C" char *utb_batch_bytes(utb_batch *b) { return b->h_bytes; }
extern "C" uint64_t *utb_batch_seq_off(utb_batch *b) { return b->h_seq_off; }
extern "C" uint32_t *utb_batch_seq_len(utb_batch *b) { return b->h_seq_len; }
extern "C" size_t utb_batch_max_bytes(const utb_batch *b) { return b->max_bytes; }
extern "C" size_t utb_batch_max_reads(const utb_batch *b) { return b->max_reads; }

extern "C" void utb_batch_destroy(utb_batch *b) {
    if (!b) return;
    cudaSetDevice(b->db->device);
    if (b->st) cudaStreamSynchronize(b->st);
    cudaFreeHost(b->h_bytes); cudaFreeHost(b->h_seq_off); cudaFreeHost(b->h_seq_len); cudaFreeHost(b->h_grp_off);
    cudaFreeHost(b->h_results); cudaFreeHost(b->h_counters);
    cudaFreeHost(b->h_name_off); cudaFreeHost(b->h_name_len); cudaFreeHost(b->h_text); cudaFreeHost(b->h_text_len);
    cudaFree(b->d_name_off); cudaFree(b->d_name_len); cudaFree(b->d_line_len); cudaFree(b->d_line_off); cudaFree(b->d_scan_sums);
    cudaFree(b->d_text_len); cudaFree(b->d_text);
    if (b->text_len_ready) cudaEventDestroy(b->text_len_ready);
    cudaFree(b->d_raw); cudaFree(b->d_seq_off); cudaFree(b->d_seq_len); cudaFree(b->d_grp_off);
    cudaFree(b->d_pk); cudaFree(b->d_bad); cudaFree(b->d_hits); cudaFree(b->d_results);
    cudaFree(b->d_gen_list); cudaFree(b->d_gen_count); cudaFree(b->d_warp_list); cudaFree(b->d_warp_count); cudaFree(b->d_counters);
    cudaFree(b->d_hist); cudaFree(b->d_tlab); cudaFree(b->d_tcnt);
    cudaFree(b->d_qwords); cudaFree(b->d_qslots); cudaFree(b->d_qcount); cudaFree(b->d_hitmap);
    cudaFree(b->d_pwords); cudaFree(b->d_ppos); cudaFree(b->d_pfill); cudaFree(b->d_pctr);
    cudaFree(b->d_nl); cudaFree(b->d_blk); cudaFree(b->d_frame_err); cudaFree(b->d_ngroups); cudaFreeHost(b->h_frame_err);
    cudaFree(b->d_frame_info); cudaFreeHost(b->h_frame_info); if (b->count_ready) cudaEventDestroy(b->count_ready);
    if (b->done) cudaEventDestroy(b->done);
    for (int i = 0; i < 7; ++i) if (b->ev[i]) cudaEventDestroy(b->ev[i]);
    if (b->st) cudaStreamDestroy(b->st);
    free(b);
}

extern "C" int utb_batch_create(utb_db *db, size_t max_bytes, size_t max_reads, utb_batch **out) {
    if (!db || !out || !max_bytes || !max_reads) { utb_set_error("utb_batch_create: bad argument"); return UTB_ERR_ARG; }
    *out = nullptr;
    if (max_bytes + 32 * max_reads >= ((uint64_t)1 << 32) - 4096) {
        utb_set_error("batch too large: positions must fit 32 bits"); return UTB_ERR_LIMIT;
    }
    CK(cudaSetDevice(db->device));
    utb_batch *b = (utb_batch *)calloc(1, sizeof(utb_batch));
    if (!b) { utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    b->db = db; b->max_bytes = max_bytes; b->max_reads = max_reads;
    // every read owns ceil((len+1)/32) <= len/32 + 1 groups
    b->max_groups = max_bytes / 32 + max_reads + 1;
    size_t npos = (size_t)b->max_groups * 32;
    size_t nl = db->d.max_ix ? db->d.max_ix : 1;
#define BK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { utb_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); utb_batch_destroy(b); return UTB_ERR_CUDA; } } while (0)
    BK(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
    BK(cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming));
    for (int i = 0; i < 7; ++i) BK(cudaEventCreate(&b->ev[i]));
    BK(cudaMallocHost(&b->h_bytes, max_bytes + 64));
    BK(cudaMallocHost(&b->h_seq_off, max_reads * 8));
    BK(cudaMallocHost(&b->h_seq_len, max_reads * 4));
    BK(cudaMallocHost(&b->h_grp_off, (max_reads + 1) * 4));
    BK(cudaMallocHost(&b->h_results, max_reads * sizeof(utb_result)));
    BK(cudaMallocHost(&b->h_counters, 4 * COUNTER_SLOTS * 8));
    BK(cudaMallocHost(&b->h_name_off, (max_reads + 1) * 4));
    BK(cudaMallocHost(&b->h_name_len, (max_reads + 1) * 4));
    BK(cudaMallocHost(&b->h_text_len, 4));
    BK(cudaEventCreateWithFlags(&b->text_len_ready, cudaEventDisableTiming));
    BK(cudaMalloc(&b->d_name_off, (max_reads + 1) * 4));
    BK(cudaMalloc(&b->d_name_len, (max_reads + 1) * 4));
    BK(cudaMalloc(&b->d_line_len, (max_reads + 1) * 4));
    BK(cudaMalloc(&b->d_line_off, (max_reads + 1) * 4));
    BK(cudaMalloc(&b->d_scan_sums, (max_reads / SCAN_TILE + 2) * 4));
    BK(cudaMalloc(&b->d_text_len, 4));
    BK(cudaMalloc(&b->d_raw, max_bytes + 128));
    BK(cudaMalloc(&b->d_seq_off, max_reads * 8));
    BK(cudaMalloc(&b->d_seq_len, max_reads * 4));
    BK(cudaMalloc(&b->d_grp_off, (max_reads + 1) * 4));
    BK(cudaMalloc(&b->d_pk, (b->max_groups + 2) * 8));
    BK(cudaMalloc(&b->d_bad, (b->max_groups + 2) * 4));
    BK(cudaMalloc(&b->d_hits, npos * 2 * 4));
    BK(cudaMalloc(&b->d_results, max_reads * sizeof(utb_result)));
    BK(cudaMalloc(&b->d_gen_list, max_reads * 4));
    BK(cudaMalloc(&b->d_gen_count, 4));
    BK(cudaMalloc(&b->d_warp_list, max_reads * 4));
    BK(cudaMalloc(&b->d_warp_count, 4));
    BK(cudaMalloc(&b->d_counters, 4 * COUNTER_SLOTS * 8));
    BK(cudaMalloc(&b->d_hist, (size_t)VB_BLOCKS * nl * 4));
    BK(cudaMalloc(&b->d_tlab, (size_t)VB_BLOCKS * nl * 4));
    BK(cudaMalloc(&b->d_tcnt, (size_t)VB_BLOCKS * nl * 4));
    BK(cudaMemset(b->d_hist, 0, (size_t)VB_BLOCKS * nl * 4));
    BK(cudaMemset(b->d_raw, 0, max_bytes + 128));
    if (db->bloom) {                                               // queue for a quarter of the lookup slots; overflow resolves inline
        b->q_cap = npos * 2 / 4 + 4096;
        BK(cudaMalloc(&b->d_qwords, b->q_cap * 8));
        BK(cudaMalloc(&b->d_qslots, b->q_cap * 4));
        BK(cudaMalloc(&b->d_qcount, 8));
        BK(cudaMalloc(&b->d_hitmap, (npos * 2 / 32 + 8) * 4));
        // Partitioned filter pass: pays once a batch holds enough positions for the probes of a partition to
        // reuse its filter lines (and to amortise 64 grid barriers); UTB_PARTITION=1 forces it, =0 disables it.
        const char *pe = getenv("UTB_PARTITION");
        b->part_min_pos = pe ? (atoi(pe) ? 0 : ~0ull) : PART_MIN_POS;
        const char *pm = getenv("UTB_PARTITION_MIN_MPOS");         // threshold in Mi positions (tuning)
        if (pm && atoi(pm) > 0) b->part_min_pos = (uint64_t)atoi(pm) << 20;
        if (npos >= b->part_min_pos) {                              // one region per (CTA, partition), 25 % head-room
            b->p_total = (uint64_t)((double)npos * 1.25) + (uint64_t)P_CTAS * NPART * 68;
            BK(cudaMalloc(&b->d_pwords, b->p_total * 8));
            BK(cudaMalloc(&b->d_ppos, b->p_total * 4));
            BK(cudaMalloc(&b->d_pfill, (size_t)P_CTAS * NPART * 4));
            BK(cudaMalloc(&b->d_pctr, NPART * 4));
        }
    }
    if (db->l2_window) {
        cudaStreamAttrValue a;
        memset(&a, 0, sizeof a);
        a.accessPolicyWindow.base_ptr = db->binix;
        a.accessPolicyWindow.num_bytes = db->l2_window_bytes;
        a.accessPolicyWindow.hitRatio = db->l2_hit_ratio;
        a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        a.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(b->st, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) cudaGetLastError();
    }
#undef BK
    *out = b;
    return UTB_OK;
}

// launches the device stages on b->st; if timed, records ev[0..4] around them
static int launch_stages(utb_batch *b, bool timed) {
    const DevDB &d = b->db->d;
    const uint32_t n_reads = (uint32_t)b->n_reads, n_groups = b->n_groups;
    const uint32_t n_pos = n_groups * 32u;
    const uint32_t nstr = b->do_rc ? 2u : 1u;
    CK(cudaMemsetAsync(b->d_gen_count, 0, 4, b->st));
    CK(cudaMemsetAsync(b->d_counters, 0, 4 * COUNTER_SLOTS * 8, b->st));
    if (timed) CK(cudaEventRecord(b->ev[0], b->st));
    if (n_reads) {
        pack_kernel<<<(n_groups + 1 + 255) / 256, 256, 0, b->st>>>(b->d_raw, b->d_seq_off, b->d_seq_len, b->d_grp_off,
                                                                   n_reads, n_groups, b->ngroups_dev, b->d_pk, b->d_bad);
        b->launches++;
    }
    if (timed) CK(cudaEventRecord(b->ev[1], b->st));
    if (n_pos) {
        const unsigned nb = (unsigned)(((uint64_t)n_pos * nstr + 255) / 256);
        // pre-filter on while misses dominate (it only adds a fetch to lookups that hit)
        const bool bloom = b->db->bloom && (b->db->bloom_mode == 1 || (b->db->bloom_mode == 2 && b->db->ema_hit_rate < 0.40));
        b->used_bloom = bloom;
        b->used_partition = 0;
        if (b->db->use_interp && bloom) {
            CK(cudaMemsetAsync(b->d_qcount, 0, 8, b->st));
            // hits are sparse (only filter survivors that really match): the survivor kernel flags them in a
            // 1-bit-per-slot map, the vote reads the map and gathers only flagged slots
            CK(cudaMemsetAsync(b->d_hitmap, 0, ((size_t)n_pos * nstr / 32 + 4) * 4, b->st));
            if (timed) CK(cudaEventRecord(b->ev[4], b->st));
            const unsigned wb = (n_pos + 255u * FILT_ILP) / (256u * FILT_ILP);   // CTAs that cover every position once
            const unsigned pb = wb < 148u * FILT_MINB ? wb : 148u * FILT_MINB;  // persistent: FILT_MINB CTAs per SM
            if (b->d_pwords && (uint64_t)n_pos >= b->part_min_pos) {
                const unsigned tiles = (unsigned)(((uint64_t)n_pos + P_TILE - 1) / P_TILE);
                const unsigned gb = tiles < P_CTAS ? tiles : P_CTAS;
                uint32_t cap_cp = ((uint32_t)((double)n_pos / ((double)gb * NPART) * 1.25) + 66u) & ~1u;   // even; <= p_total / (gb * NPART)
                // measured on B200 (10 M reads): atomic ranks 11.5 ms, match ranks 16.2 ms -> atomics by default
                const char *pa = getenv("UTB_PART_ATOMS");                  // tests: 0 selects the match-ranked variant
                const int part_atoms = !(pa && atoi(pa) == 0);
#define LAUNCH_PART(NS, M, SM) do { \
                    CK(cudaFuncSetAttribute(partition_kernel<NS, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); \
                    partition_kernel<NS, M><<<gb, 256, SM, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_counters, b->d_pwords, b->d_ppos, b->d_pfill, cap_cp, b->d_hits, b->d_hitmap); } while (0)
                if (nstr == 2) { if (part_atoms) LAUNCH_PART(2, false, P_SMEM); else LAUNCH_PART(2, true, P_SMEM_MATCH); }
                else { if (part_atoms) LAUNCH_PART(1, false, P_SMEM); else LAUNCH_PART(1, true, P_SMEM_MATCH); }
#undef LAUNCH_PART
                if (timed) CK(cudaEventRecord(b->ev[6], b->st));
                b->used_partition = 1;
                CK(cudaMemsetAsync(b->d_pctr, 0, NPART * 4, b->st));
                // cooperative launch: the grid barrier between partitions needs every CTA resident
                const void *pk_fn = nstr == 2 ? (const void *)probe_kernel<2> : (const void *)probe_kernel<1>;
                int per_sm = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk_fn, 256, 0));
                if (per_sm > 4) per_sm = 4;
                int sms = 148;                                      // B200; a co-resident grid must not assume more than the device has
                if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->db->device) != cudaSuccess || sms < 1) { cudaGetLastError(); sms = 148; }
                unsigned pgrid = (unsigned)sms * (unsigned)(per_sm > 0 ? per_sm : 1), n_ctas = gb;
                DevDB dd = d;
                void *args[] = {&dd, &b->d_pwords, &b->d_ppos, &b->d_pfill, &cap_cp, &n_ctas, &b->d_pctr, &b->d_qwords, &b->d_qslots,
                                &b->d_qcount, &b->q_cap, &b->d_hits, &b->d_hitmap, &b->d_counters};
                CK(cudaLaunchCooperativeKernel(pk_fn, dim3(pgrid), dim3(256), args, 0, b->st));
                b->launches++;
            } else if (nstr == 2) filter_kernel<2><<<pb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters, b->d_qwords, b->d_qslots, b->d_qcount, b->q_cap, b->d_hitmap);
            else filter_kernel<1><<<pb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters, b->d_qwords, b->d_qslots, b->d_qcount, b->q_cap, b->d_hitmap);
            if (timed) CK(cudaEventRecord(b->ev[5], b->st));
            queue_lookup_kernel<<<148 * 6, 256, 0, b->st>>>(d, b->d_qwords, b->d_qslots, b->d_qcount, b->q_cap, b->d_hits, b->d_hitmap, b->d_counters);
            b->launches++;
        } else if (b->db->use_interp) {
            if (nstr == 2) lookup_kernel<2, true, false><<<nb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters);
            else lookup_kernel<1, true, false><<<nb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters);
        } else {
            if (nstr == 2) lookup_kernel<2, false, false><<<nb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters);
            else lookup_kernel<1, false, false><<<nb, 256, 0, b->st>>>(d, b->d_pk, b->d_bad, n_pos, b->ngroups_dev, b->d_hits, b->d_counters);
        }
        b->launches++;
    }
    if (timed) CK(cudaEventRecord(b->ev[2], b->st));
    if (n_reads) {
        VoteIn in;
        in.hits = b->d_hits; in.grp_off = b->d_grp_off; in.seq_len = b->d_seq_len; in.off = nullptr; in.nstr = nstr;
        in.hitmap = b->used_bloom ? b->d_hitmap : nullptr;
        if (in.hitmap) {
            // thread per read for the common case; label-rich or long reads fall through to the warp kernel
            // (and from there to the block kernel) by way of device-side lists
            CK(cudaMemsetAsync(b->d_warp_count, 0, 4, b->st));
            vote_thread_kernel<<<(n_reads + VT_THREADS - 1) / VT_THREADS, VT_THREADS, 0, b->st>>>(
                d, in, n_reads, b->d_results, b->d_warp_list, b->d_warp_count, b->d_counters);
            vote_warp_kernel<<<148 * 6, VW_WARPS * 32, 0, b->st>>>(
                d, in, n_reads, b->d_warp_list, b->d_warp_count, b->d_results, b->d_gen_list, b->d_gen_count, b->d_counters);
            b->launches++;
        } else
            vote_warp_kernel<<<(n_reads + VW_WARPS - 1) / VW_WARPS, VW_WARPS * 32, 0, b->st>>>(
                d, in, n_reads, nullptr, nullptr, b->d_results, b->d_gen_list, b->d_gen_count, b->d_counters);
        vote_block_kernel<<<VB_BLOCKS, VB_THREADS, 0, b->st>>>(d, in, b->d_results, b->d_gen_list, b->d_gen_count,
                                                               b->d_hist, b->d_tlab, b->d_tcnt, b->d_counters);
        b->launches += 2;
    }
    if (timed) CK(cudaEventRecord(b->ev[3], b->st));
    CK(cudaGetLastError());
    return UTB_OK;
}

// text buffers are allocated on first use (the device-resident measurement batches never format)
static int ensure_text_buffers(utb_batch *b) {
    if (b->d_text) return UTB_OK;
    size_t cap = b->max_bytes + b->max_reads * (b->db->max_label + 48) + 64;
    if (cap >= ((size_t)1 << 32)) { utb_set_error("batch too large for device-side formatting"); return UTB_ERR_LIMIT; }
    CK(cudaMalloc(&b->d_text, cap));
    CK(cudaMallocHost(&b->h_text, cap));
    b->text_cap = cap;
    return UTB_OK;
}

static int submit_impl(utb_batch *b, const char *src, size_t n_bytes, size_t n_reads, int do_rc, int want_text, uint64_t total_groups);
static int finish_submit(utb_batch *b);
extern "C" int utb_batch_submit(utb_batch *b, size_t n_bytes, size_t n_reads, int do_rc) {
    return submit_impl(b, nullptr, n_bytes, n_reads, do_rc, 0, 0);
}
extern "C" int utb_batch_submit_text(utb_batch *b, size_t n_bytes, size_t n_reads, int do_rc) {
    return submit_impl(b, nullptr, n_bytes, n_reads, do_rc, 1, 0);
}
// Pipeline-internal variant: `src` (pinned host memory, or NULL for the batch's own staging) holds the raw
// bytes; total_groups != 0 means the caller already summed utb_read_slots() over the reads, so the
// per-read offsets are scanned on the device instead of in a serial host loop.
extern "C" int utb_batch_submit_ex(utb_batch *b, const char *src, size_t n_bytes, size_t n_reads, int do_rc,
                                   int want_text, uint64_t total_groups) {
    return submit_impl(b, src, n_bytes, n_reads, do_rc, want_text, total_groups);
}
extern "C" int utb_host_ptr_is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost;
}
extern "C" uint32_t *utb_batch_name_off(utb_batch *b) { return b->h_name_off; }
extern "C" uint32_t *utb_batch_name_len(utb_batch *b) { return b->h_name_len; }

static int submit_impl(utb_batch *b, const char *src, size_t n_bytes, size_t n_reads, int do_rc, int want_text, uint64_t total_groups) {
    if (!b) { utb_set_error("utb_batch_submit: null batch"); return UTB_ERR_ARG; }
    if (n_bytes > b->max_bytes || n_reads > b->max_reads) { utb_set_error("utb_batch_submit: batch over capacity"); return UTB_ERR_LIMIT; }
    CK(cudaSetDevice(b->db->device));
    // position space: read r owns groups [grp_off[r], grp_off[r+1])
    uint64_t g = total_groups;
    if (!total_groups) for (size_t r = 0; r < n_reads; ++r) {
        uint32_t len = b->h_seq_len[r];
        if (len > UTB_MAXSEQ) { utb_set_error("read %zu: %u bases exceeds the 16777214-base limit", r, len); return UTB_ERR_LIMIT; }
        if (b->h_seq_off[r] + len > n_bytes) { utb_set_error("read %zu: sequence outside the batch bytes", r); return UTB_ERR_ARG; }
        b->h_grp_off[r] = (uint32_t)g;
        g += utb_read_slots(len);
    }
    if (g > b->max_groups) { utb_set_error("utb_batch_submit: %llu position groups exceed capacity %llu", (unsigned long long)g, (unsigned long long)b->max_groups); return UTB_ERR_LIMIT; }
    if (!total_groups) b->h_grp_off[n_reads] = (uint32_t)g;
    b->n_reads = n_reads; b->n_groups = (uint32_t)g; b->do_rc = do_rc ? 1 : 0;
    b->want_text = want_text; b->ngroups_dev = nullptr; b->framed_on_device = 0;
    if (want_text) {
        int rt = ensure_text_buffers(b);
        if (rt) return rt;
        if (n_reads) {
            CK(cudaMemcpyAsync(b->d_name_off, b->h_name_off, n_reads * 4, cudaMemcpyHostToDevice, b->st));
            CK(cudaMemcpyAsync(b->d_name_len, b->h_name_len, n_reads * 4, cudaMemcpyHostToDevice, b->st));
        }
    }
    if (n_reads) {
        CK(cudaMemcpyAsync(b->d_raw, src ? src : b->h_bytes, n_bytes, cudaMemcpyHostToDevice, b->st));
        CK(cudaMemcpyAsync(b->d_seq_off, b->h_seq_off, n_reads * 8, cudaMemcpyHostToDevice, b->st));
        CK(cudaMemcpyAsync(b->d_seq_len, b->h_seq_len, n_reads * 4, cudaMemcpyHostToDevice, b->st));
        if (!total_groups) CK(cudaMemcpyAsync(b->d_grp_off, b->h_grp_off, (n_reads + 1) * 4, cudaMemcpyHostToDevice, b->st));
        else {                                                     // grp_off = exclusive scan of the per-read slot counts
            const uint32_t n = (uint32_t)n_reads, nb = (n + SCAN_TILE - 1) / SCAN_TILE;
            slots_kernel<<<(n + 255) / 256, 256, 0, b->st>>>(b->d_seq_len, n, b->d_line_len);
            scan_sums_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums);
            scan_top_kernel<<<1, 1024, 0, b->st>>>(b->d_scan_sums, nb, b->d_text_len);
            scan_apply_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums, b->d_grp_off);
            b->launches += 4;
        }
    }
    return finish_submit(b);
}

// device stages + (text | result records) + counters back to the host, all on the batch's stream
static int finish_submit(utb_batch *b) {
    const size_t n_reads = b->n_reads;
    const int want_text = b->want_text;
    int rc = launch_stages(b, true);
    if (rc) return rc;
    if (want_text) {
        // lines built on the device: lengths -> exclusive scan -> one warp per read writes its line
        const uint32_t n = (uint32_t)n_reads, nb = (n + SCAN_TILE - 1) / SCAN_TILE;
        CK(cudaMemsetAsync(b->d_text_len, 0, 4, b->st));
        if (n) {
            fmt_len_kernel<<<(n + 255) / 256, 256, 0, b->st>>>(b->db->d, b->d_results, b->d_name_len, n, b->d_line_len);
            scan_sums_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums);
            scan_top_kernel<<<1, 1024, 0, b->st>>>(b->d_scan_sums, nb, b->d_text_len);
            scan_apply_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums, b->d_line_off);
            fmt_write_kernel<<<(n + FW_WARPS - 1) / FW_WARPS, FW_WARPS * 32, 0, b->st>>>(
                b->db->d, b->d_results, b->d_raw, b->d_name_off, b->d_name_len, b->d_line_off, n, b->d_text);
            b->launches += 5;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(b->h_text_len, b->d_text_len, 4, cudaMemcpyDeviceToHost, b->st));
        // the text length is only known on the device: copy an estimate now (no extra round trip in the
        // common case), wait_text tops it up if it was short
        size_t est = (size_t)(b->db->text_per_read * 1.15 * (double)n_reads) + 4096;
        if (b->db->text_per_read <= 0) est = 0;
        if (est > b->text_cap) est = b->text_cap;
        b->text_prefetched = est;
        if (est) CK(cudaMemcpyAsync(b->h_text, b->d_text, est, cudaMemcpyDeviceToHost, b->st));
    } else if (n_reads) CK(cudaMemcpyAsync(b->h_results, b->d_results, n_reads * sizeof(utb_result), cudaMemcpyDeviceToHost, b->st));
    CK(cudaMemcpyAsync(b->h_counters, b->d_counters, 4 * COUNTER_SLOTS * 8, cudaMemcpyDeviceToHost, b->st));
    CK(cudaEventRecord(b->done, b->st));
    b->in_flight = 1;
    return UTB_OK;
}

// Raw submit (pipeline-internal), in two steps so that the host need not look at the bytes at all:
//   utb_batch_raw_begin  copies the chunk to the device and counts its newlines there (and notes a NUL byte);
//   utb_batch_raw_count  blocks until that count is on the host (only this batch's copy and one small kernel
//                        are waited for; the other stream slots keep the GPU busy meanwhile);
//   utb_batch_raw_finish the caller has derived from the count that the first n_bytes hold exactly n_reads
//                        complete records = 2 * n_reads lines, each ended by '\n', and no NUL byte: the records
//                        are framed on the device, searched, and the output text is built there.
// utb_batch_submit_raw = begin + finish for a caller that counted the newlines itself.
// utb_batch_frame_error() tells after the wait whether a record was malformed.
extern "C" int utb_batch_raw_begin(utb_batch *b, const char *src, size_t n_bytes) {
    if (!b || !n_bytes) { utb_set_error("utb_batch_raw_begin: bad argument"); return UTB_ERR_ARG; }
    if (n_bytes > b->max_bytes) { utb_set_error("utb_batch_raw_begin: batch over capacity"); return UTB_ERR_LIMIT; }
    CK(cudaSetDevice(b->db->device));
    if (!b->d_nl) {
        CK(cudaMalloc(&b->d_nl, (2 * b->max_reads + 2) * 4));
        CK(cudaMalloc(&b->d_blk, (b->max_bytes / FR_BLOCK + 2) * 4));
        CK(cudaMalloc(&b->d_frame_err, 4));
        CK(cudaMalloc(&b->d_ngroups, 4));
        CK(cudaMalloc(&b->d_frame_info, 8));
        CK(cudaMallocHost(&b->h_frame_err, 4));
        CK(cudaMallocHost(&b->h_frame_info, 8));
        CK(cudaEventCreateWithFlags(&b->count_ready, cudaEventDisableTiming));
    }
    int rt = ensure_text_buffers(b);
    if (rt) return rt;
    b->raw_bytes = n_bytes;
    CK(cudaMemcpyAsync(b->d_raw, src ? src : b->h_bytes, n_bytes, cudaMemcpyHostToDevice, b->st));
    CK(cudaMemsetAsync(b->d_frame_err, 0xFF, 4, b->st));
    CK(cudaMemsetAsync(b->d_frame_info, 0, 8, b->st));
    const uint32_t fb = (uint32_t)((n_bytes + FR_BLOCK - 1) / FR_BLOCK);
    nl_count_kernel<<<fb, 256, 0, b->st>>>(b->d_raw, n_bytes, b->d_blk, b->d_frame_info + 1);
    scan_top_kernel<<<1, 1024, 0, b->st>>>(b->d_blk, fb, b->d_frame_info);       // block counts -> exclusive offsets, total newlines
    b->launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(b->h_frame_info, b->d_frame_info, 8, cudaMemcpyDeviceToHost, b->st));
    CK(cudaEventRecord(b->count_ready, b->st));
    return UTB_OK;
}
extern "C" int utb_batch_raw_count(utb_batch *b, size_t *n_newlines, int *has_nul) {
    if (!b || !b->count_ready || !n_newlines || !has_nul) { utb_set_error("utb_batch_raw_count: bad argument"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    CK(cudaEventSynchronize(b->count_ready));
    *n_newlines = b->h_frame_info[0]; *has_nul = b->h_frame_info[1] != 0;
    return UTB_OK;
}
extern "C" int utb_batch_raw_finish(utb_batch *b, size_t n_bytes, size_t n_reads, int do_rc) {
    if (!b || !n_reads || !n_bytes || n_bytes > b->raw_bytes) { utb_set_error("utb_batch_raw_finish: bad argument"); return UTB_ERR_ARG; }
    if (n_reads > b->max_reads) { utb_set_error("utb_batch_raw_finish: batch over capacity"); return UTB_ERR_LIMIT; }
    CK(cudaSetDevice(b->db->device));
    // every read owns ceil((len+1)/32) <= len/32 + 1 groups and its two lines hold at least 2 more bytes than its bases
    uint64_t g_ub = (n_bytes - 2 * n_reads) / 32 + n_reads + 1;
    if (n_bytes < 2 * n_reads || g_ub > b->max_groups) g_ub = b->max_groups;
    b->n_reads = n_reads; b->n_groups = (uint32_t)g_ub; b->do_rc = do_rc ? 1 : 0;
    b->want_text = 1; b->ngroups_dev = b->d_ngroups; b->framed_on_device = 1;
    *b->h_frame_err = FR_ERR_NONE;
    // the block offsets were scanned over the whole chunk; newlines at or beyond n_bytes have index >= 2 * n_reads
    const uint32_t n = (uint32_t)n_reads, fb = (uint32_t)((b->raw_bytes + FR_BLOCK - 1) / FR_BLOCK), nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    nl_index_kernel<<<fb, 256, 0, b->st>>>(b->d_raw, n_bytes, b->d_blk, 2 * n, b->d_nl);
    frame_parse_kernel<<<(n + 255) / 256, 256, 0, b->st>>>(b->d_raw, b->d_nl, n, b->d_seq_off, b->d_seq_len, b->d_name_off, b->d_name_len,
                                                          b->d_line_len, b->d_frame_err);
    // grp_off = exclusive scan of the per-read group counts; the total stays on the device
    scan_sums_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums);
    scan_top_kernel<<<1, 1024, 0, b->st>>>(b->d_scan_sums, nb, b->d_ngroups);
    scan_apply_kernel<<<nb, 256, 0, b->st>>>(b->d_line_len, n, b->d_scan_sums, b->d_grp_off);
    b->launches += 5;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(b->h_frame_err, b->d_frame_err, 4, cudaMemcpyDeviceToHost, b->st));
    return finish_submit(b);
}
extern "C" int utb_batch_submit_raw(utb_batch *b, const char *src, size_t n_bytes, size_t n_reads, int do_rc) {
    int rc = utb_batch_raw_begin(b, src, n_bytes);
    return rc ? rc : utb_batch_raw_finish(b, n_bytes, n_reads, do_rc);
}
// After the wait of a raw submit: 0 if every record was well formed, else 1 with the index of the first
// malformed record of the batch and its FE_* code (1 no header '>', 2 sequence begins '>', 4 line too long).
extern "C" int utb_batch_frame_error(const utb_batch *b, size_t *record, int *code) {
    if (!b || !b->framed_on_device || !b->h_frame_err || *b->h_frame_err == FR_ERR_NONE) return 0;
    if (record) *record = *b->h_frame_err >> 3;
    if (code) *code = (int)(*b->h_frame_err & 7u);
    return 1;
}

extern "C" int utb_batch_wait(utb_batch *b, const utb_result **results) {
    if (!b) { utb_set_error("utb_batch_wait: null batch"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    CK(cudaEventSynchronize(b->done));
    if (b->in_flight) {
        uint64_t l = 0, h = 0;
        for (int i = 0; i < COUNTER_SLOTS; ++i) { l += b->h_counters[i]; h += b->h_counters[COUNTER_SLOTS + i]; }
        if (l >= 4096) {                                            // steer the pre-filter: hit rate seen so far
            double r = (double)h / (double)l;
            b->db->ema_hit_rate = b->db->ema_hit_rate < 0 ? r : 0.5 * b->db->ema_hit_rate + 0.5 * r;
        }
    }
    b->in_flight = 0;
    if (results) *results = b->h_results;
    return UTB_OK;
}

// After utb_batch_submit_text: blocks until the batch is done and its output text is in pinned host memory.
extern "C" int utb_batch_wait_text(utb_batch *b, const char **text, size_t *len, uint64_t *good_finds) {
    if (!b || !text || !len) { utb_set_error("utb_batch_wait_text: null argument"); return UTB_ERR_ARG; }
    if (!b->want_text) { utb_set_error("utb_batch_wait_text: batch was not submitted with utb_batch_submit_text"); return UTB_ERR_ARG; }
    int rc = utb_batch_wait(b, nullptr);
    if (rc) return rc;
    const size_t n = *b->h_text_len;
    if (n > b->text_cap) { utb_set_error("device text overflow"); return UTB_ERR_LIMIT; }
    if (n > b->text_prefetched) {
        CK(cudaMemcpyAsync(b->h_text + b->text_prefetched, b->d_text + b->text_prefetched, n - b->text_prefetched,
                           cudaMemcpyDeviceToHost, b->st));
        CK(cudaStreamSynchronize(b->st));
    }
    if (b->n_reads >= 64) {
        double r = (double)n / (double)b->n_reads;
        b->db->text_per_read = b->db->text_per_read <= 0 ? r : 0.5 * b->db->text_per_read + 0.5 * r;
    }
    *text = b->h_text; *len = n;
    if (good_finds) { uint64_t g = 0; for (int i = 0; i < COUNTER_SLOTS; ++i) g += b->h_counters[2 * COUNTER_SLOTS + i]; *good_finds = g; }
    return UTB_OK;
}

extern "C" int utb_batch_counts(utb_batch *b, uint64_t *lookups, uint64_t *hits) {
    if (!b) { utb_set_error("utb_batch_counts: null batch"); return UTB_ERR_ARG; }
    uint64_t l = 0, h = 0;
    for (int i = 0; i < COUNTER_SLOTS; ++i) { l += b->h_counters[i]; h += b->h_counters[COUNTER_SLOTS + i]; }
    if (lookups) *lookups = l;
    if (hits) *hits = h;
    return UTB_OK;
}

// device-stage milliseconds of the LAST submit (valid after wait): pack, lookup, vote, total
extern "C" int utb_batch_last_ms(utb_batch *b, float ms[4]) {
    if (!b || !ms) { utb_set_error("utb_batch_last_ms: null argument"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    for (int i = 0; i < 3; ++i) CK(cudaEventElapsedTime(&ms[i], b->ev[i], b->ev[i + 1]));
    CK(cudaEventElapsedTime(&ms[3], b->ev[0], b->ev[3]));
    return UTB_OK;
}
extern "C" uint64_t utb_batch_launches(const utb_batch *b) { return b ? b->launches : 0; }
// Two-phase detail of the LAST run (valid after wait): ms[0] filter kernel, ms[1] queue kernel (0/0 when the
// single lookup kernel ran); sectors[0] = lookups answered by the filter (one sector per position, both
// strands), sectors[1] = sectors the exact path touched for the survivors.
// ms[0] partition_kernel, ms[1] probe_kernel of the LAST run; both 0 when the direct filter kernel ran.
extern "C" int utb_batch_partition_detail(utb_batch *b, float ms[2]) {
    if (!b || !ms) { utb_set_error("utb_batch_partition_detail: null argument"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    ms[0] = ms[1] = 0;
    if (!b->used_bloom || !b->used_partition) return UTB_OK;
    CK(cudaEventElapsedTime(&ms[0], b->ev[4], b->ev[6]));
    CK(cudaEventElapsedTime(&ms[1], b->ev[6], b->ev[5]));
    return UTB_OK;
}
extern "C" int utb_batch_lookup_detail(utb_batch *b, float ms[2], uint64_t sectors[2]) {
    if (!b || !ms || !sectors) { utb_set_error("utb_batch_lookup_detail: null argument"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    ms[0] = ms[1] = 0; sectors[0] = sectors[1] = 0;
    if (!b->used_bloom) return UTB_OK;
    CK(cudaEventElapsedTime(&ms[0], b->ev[4], b->ev[5]));   // the kernel alone (the hits fill before it is part of the stage)
    CK(cudaEventElapsedTime(&ms[1], b->ev[5], b->ev[2]));
    for (int i = 0; i < COUNTER_SLOTS; ++i) { sectors[0] += b->h_counters[i]; sectors[1] += b->h_counters[3 * COUNTER_SLOTS + i]; }
    return UTB_OK;
}

extern "C" int utb_batch_rerun_device(utb_batch *b, int iters, float ms[4], uint64_t *launches) {
    if (!b || iters < 1) { utb_set_error("utb_batch_rerun_device: bad argument"); return UTB_ERR_ARG; }
    CK(cudaSetDevice(b->db->device));
    CK(cudaStreamSynchronize(b->st));
    float acc[4] = {0, 0, 0, 0};
    uint64_t l0 = b->launches;
    for (int it = 0; it < iters; ++it) {
        int rc = launch_stages(b, true);
        if (rc) return rc;
        CK(cudaMemcpyAsync(b->h_counters, b->d_counters, 4 * COUNTER_SLOTS * 8, cudaMemcpyDeviceToHost, b->st));
        CK(cudaStreamSynchronize(b->st));
        float t[4];
        rc = utb_batch_last_ms(b, t);
        if (rc) return rc;
        for (int i = 0; i < 4; ++i) acc[i] += t[i];
    }
    if (ms) for (int i = 0; i < 4; ++i) ms[i] = acc[i];
    if (launches) *launches = b->launches - l0;
    return UTB_OK;
}

// ---------------------------------------------------------------------------
// C ABI: stage-level entry points for parity tests
// ---------------------------------------------------------------------------
extern "C" int utb_lookup_words(utb_db *db, const uint64_t *words, size_t n, uint32_t *ix) {
    if (!db || (!words && n) || (!ix && n)) { utb_set_error("utb_lookup_words: null argument"); return UTB_ERR_ARG; }
    if (!n) return UTB_OK;
    CK(cudaSetDevice(db->device));
    uint64_t *dw = nullptr; uint32_t *di = nullptr;
    CK(cudaMalloc(&dw, n * 8));
    CK(cudaMalloc(&di, n * 4));
    CK(cudaMemcpy(dw, words, n * 8, cudaMemcpyHostToDevice));
    if (db->use_interp) lookup_words_kernel<true><<<(unsigned)((n + 255) / 256), 256>>>(db->d, dw, n, di);
    else lookup_words_kernel<false><<<(unsigned)((n + 255) / 256), 256>>>(db->d, dw, n, di);
    CK(cudaGetLastError());
    CK(cudaMemcpy(ix, di, n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dw); cudaFree(di);
    return UTB_OK;
}

extern "C" int utb_pack_sequence(utb_db *db, const char *seq, uint32_t len, uint64_t *fwd, uint64_t *rc, uint8_t *valid) {
    if (!db || !seq || !fwd || !rc || !valid) { utb_set_error("utb_pack_sequence: null argument"); return UTB_ERR_ARG; }
    if (len > UTB_MAXSEQ) { utb_set_error("sequence too long"); return UTB_ERR_LIMIT; }
    if (!len) return UTB_OK;
    CK(cudaSetDevice(db->device));
    // deliberately misaligned by 1 byte, as sequences are inside a FASTA chunk
    uint32_t n_groups = (uint32_t)utb_read_slots(len), n_pos = n_groups * 32;
    uint8_t *d_raw; uint64_t *d_off, *d_pk, *d_f, *d_r; uint32_t *d_len, *d_grp, *d_bad; uint8_t *d_v;
    CK(cudaMalloc(&d_raw, (size_t)len + 128)); CK(cudaMemset(d_raw, 0, (size_t)len + 128));
    CK(cudaMalloc(&d_off, 8)); CK(cudaMalloc(&d_len, 4)); CK(cudaMalloc(&d_grp, 8));
    CK(cudaMalloc(&d_pk, (size_t)(n_groups + 2) * 8)); CK(cudaMalloc(&d_bad, (size_t)(n_groups + 2) * 4));
    CK(cudaMalloc(&d_f, (size_t)n_pos * 8)); CK(cudaMalloc(&d_r, (size_t)n_pos * 8)); CK(cudaMalloc(&d_v, n_pos));
    uint64_t off = 1; uint32_t grp[2] = {0, n_groups};
    CK(cudaMemcpy(d_raw + 1, seq, len, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, &off, 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_len, &len, 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_grp, grp, 8, cudaMemcpyHostToDevice));
    pack_kernel<<<(n_groups + 1 + 255) / 256, 256>>>(d_raw, d_off, d_len, d_grp, 1, n_groups, nullptr, d_pk, d_bad);
    expand_windows_kernel<<<(n_pos + 255) / 256, 256>>>(d_pk, d_bad, n_pos, d_f, d_r, d_v);
    CK(cudaGetLastError());
    CK(cudaMemcpy(fwd, d_f, (size_t)len * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rc, d_r, (size_t)len * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(valid, d_v, len, cudaMemcpyDeviceToHost));
    cudaFree(d_raw); cudaFree(d_off); cudaFree(d_len); cudaFree(d_grp); cudaFree(d_pk); cudaFree(d_bad);
    cudaFree(d_f); cudaFree(d_r); cudaFree(d_v);
    return UTB_OK;
}

extern "C" int utb_vote_hits(utb_db *db, const uint32_t *hits, const uint64_t *off, size_t n_reads, utb_result *results) {
    if (!db || !off || !results || (!hits && n_reads && off[n_reads])) { utb_set_error("utb_vote_hits: null argument"); return UTB_ERR_ARG; }
    if (!n_reads) return UTB_OK;
    CK(cudaSetDevice(db->device));
    size_t nh = off[n_reads], nl = db->d.max_ix ? db->d.max_ix : 1;
    uint32_t *d_hits, *d_gl, *d_gc, *d_hist, *d_tl, *d_tc; uint64_t *d_off; utb_result *d_res; unsigned long long *d_cnt;
    CK(cudaMalloc(&d_hits, (nh + 1) * 4)); CK(cudaMalloc(&d_off, (n_reads + 1) * 8));
    CK(cudaMalloc(&d_res, n_reads * sizeof(utb_result)));
    CK(cudaMalloc(&d_gl, n_reads * 4)); CK(cudaMalloc(&d_gc, 4)); CK(cudaMalloc(&d_cnt, 4 * COUNTER_SLOTS * 8));
    CK(cudaMalloc(&d_hist, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMalloc(&d_tl, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMalloc(&d_tc, (size_t)VB_BLOCKS * nl * 4));
    CK(cudaMemset(d_hist, 0, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMemset(d_gc, 0, 4)); CK(cudaMemset(d_cnt, 0, 4 * COUNTER_SLOTS * 8));
    if (nh) CK(cudaMemcpy(d_hits, hits, nh * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, off, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    VoteIn in; in.hits = d_hits; in.grp_off = nullptr; in.seq_len = nullptr; in.off = d_off; in.nstr = 1; in.hitmap = nullptr;
    vote_warp_kernel<<<(unsigned)((n_reads + VW_WARPS - 1) / VW_WARPS), VW_WARPS * 32>>>(db->d, in, (uint32_t)n_reads, nullptr, nullptr, d_res, d_gl, d_gc, d_cnt);
    vote_block_kernel<<<VB_BLOCKS, VB_THREADS>>>(db->d, in, d_res, d_gl, d_gc, d_hist, d_tl, d_tc, d_cnt);
    CK(cudaGetLastError());
    CK(cudaMemcpy(results, d_res, n_reads * sizeof(utb_result), cudaMemcpyDeviceToHost));
    cudaFree(d_hits); cudaFree(d_off); cudaFree(d_res); cudaFree(d_gl); cudaFree(d_gc); cudaFree(d_cnt);
    cudaFree(d_hist); cudaFree(d_tl); cudaFree(d_tc);
    return UTB_OK;
}

// Same vote through the sparse representation the batch pipeline uses: every read's slots start at a
// multiple of 32, a 1-bit-per-slot hit map flags the labels, and the reads go thread kernel -> warp
// kernel -> block kernel by way of the device-side deferral lists.
extern "C" int utb_vote_hits_sparse(utb_db *db, const uint32_t *hits, const uint64_t *off, size_t n_reads, utb_result *results) {
    if (!db || !off || !results || (!hits && n_reads && off[n_reads])) { utb_set_error("utb_vote_hits_sparse: null argument"); return UTB_ERR_ARG; }
    if (!n_reads) return UTB_OK;
    CK(cudaSetDevice(db->device));
    uint64_t *poff = (uint64_t *)malloc((n_reads + 1) * 8);
    if (!poff) { utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    uint64_t tot = 0;
    for (size_t r = 0; r < n_reads; ++r) { poff[r] = tot; tot += (off[r + 1] - off[r] + 31) / 32 * 32; }
    poff[n_reads] = tot;
    uint32_t *ph = (uint32_t *)malloc((tot + 32) * 4), *pm = (uint32_t *)calloc(tot / 32 + 2, 4);
    if (!ph || !pm) { free(poff); free(ph); free(pm); utb_set_error("out of memory"); return UTB_ERR_NOMEM; }
    for (uint64_t i = 0; i < tot + 32; ++i) ph[i] = HIT_MISS;
    for (size_t r = 0; r < n_reads; ++r)
        for (uint64_t i = 0; i < off[r + 1] - off[r]; ++i) {
            const uint32_t h = hits[off[r] + i];
            ph[poff[r] + i] = h;
            if (h < HIT_NOWIN) pm[(poff[r] + i) >> 5] |= 1u << ((poff[r] + i) & 31u);
        }
    const size_t nl = db->d.max_ix ? db->d.max_ix : 1;
    uint32_t *d_hits, *d_map, *d_gl, *d_gc, *d_wl, *d_wc, *d_hist, *d_tl, *d_tc; uint64_t *d_off; utb_result *d_res; unsigned long long *d_cnt;
    CK(cudaMalloc(&d_hits, (tot + 32) * 4)); CK(cudaMalloc(&d_map, (tot / 32 + 2) * 4)); CK(cudaMalloc(&d_off, (n_reads + 1) * 8));
    CK(cudaMalloc(&d_res, n_reads * sizeof(utb_result)));
    CK(cudaMalloc(&d_gl, n_reads * 4)); CK(cudaMalloc(&d_gc, 4)); CK(cudaMalloc(&d_wl, n_reads * 4)); CK(cudaMalloc(&d_wc, 4));
    CK(cudaMalloc(&d_cnt, 4 * COUNTER_SLOTS * 8));
    CK(cudaMalloc(&d_hist, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMalloc(&d_tl, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMalloc(&d_tc, (size_t)VB_BLOCKS * nl * 4));
    CK(cudaMemset(d_hist, 0, (size_t)VB_BLOCKS * nl * 4)); CK(cudaMemset(d_gc, 0, 4)); CK(cudaMemset(d_wc, 0, 4)); CK(cudaMemset(d_cnt, 0, 4 * COUNTER_SLOTS * 8));
    CK(cudaMemcpy(d_hits, ph, (tot + 32) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_map, pm, (tot / 32 + 2) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_off, poff, (n_reads + 1) * 8, cudaMemcpyHostToDevice));
    free(poff); free(ph); free(pm);
    VoteIn in; in.hits = d_hits; in.grp_off = nullptr; in.seq_len = nullptr; in.off = d_off; in.nstr = 1; in.hitmap = d_map;
    vote_thread_kernel<<<(unsigned)((n_reads + VT_THREADS - 1) / VT_THREADS), VT_THREADS>>>(db->d, in, (uint32_t)n_reads, d_res, d_wl, d_wc, d_cnt);
    vote_warp_kernel<<<148 * 6, VW_WARPS * 32>>>(db->d, in, (uint32_t)n_reads, d_wl, d_wc, d_res, d_gl, d_gc, d_cnt);
    vote_block_kernel<<<VB_BLOCKS, VB_THREADS>>>(db->d, in, d_res, d_gl, d_gc, d_hist, d_tl, d_tc, d_cnt);
    CK(cudaGetLastError());
    CK(cudaMemcpy(results, d_res, n_reads * sizeof(utb_result), cudaMemcpyDeviceToHost));
    cudaFree(d_hits); cudaFree(d_map); cudaFree(d_off); cudaFree(d_res); cudaFree(d_gl); cudaFree(d_gc); cudaFree(d_wl); cudaFree(d_wc);
    cudaFree(d_cnt); cudaFree(d_hist); cudaFree(d_tl); cudaFree(d_tc);
    return UTB_OK;
}

// ---------------------------------------------------------------------------
// C ABI: roofline denominator
// ---------------------------------------------------------------------------
extern "C" int utb_measure_rand32(int device, uint64_t ws_bytes, uint64_t loads, int iters, double *gbs) {
    if (!gbs || ws_bytes < 4096 || iters < 1) { utb_set_error("utb_measure_rand32: bad argument"); return UTB_ERR_ARG; }
    int rc = check_device(device);
    if (rc) return rc;
    CK(cudaSetDevice(device));
    uint8_t *buf; unsigned long long *sink;
    CK(cudaMalloc(&buf, ws_bytes));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(buf, 1, ws_bytes));
    CK(cudaMemset(sink, 0, 8));
    const uint32_t per_thread = 64;
    uint64_t threads = (loads + per_thread - 1) / per_thread;
    uint32_t blocks = (uint32_t)((threads + 255) / 256);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0;
    for (int it = 0; it < iters + 1; ++it) {
        CK(cudaEventRecord(e0));
        rand32_kernel<<<blocks, 256>>>(buf, ws_bytes / 32, per_thread, 0x1234 + it, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double g = (double)blocks * 256 * per_thread * 32.0 / (ms * 1e-3) / 1e9;
        if (it > 0 && g > best) best = g;                          // first pass is warm-up
    }
    *gbs = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return UTB_OK;
}
