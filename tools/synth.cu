// synth.cu -- synthetic-universe generator for bench.py and the large parity
// tests (libutb_synth.so).  NOT part of the search path and never linked into
// libutree_b200.so: it only manufactures inputs -- a CTR file of the shape the
// reference builder + compressor would emit and a
// FASTA of reads sampled from the same synthetic genomes.
//
// Universe (pure function of the seed): n_phyla x n_genera x n_species x
// n_strains genomes of genome_len bases; a strain is its species ancestor with
// 0.5 % point mutations, a species its genus ancestor with 3 %, a genus its
// phylum ancestor with 10 % (App. E).  Bases are never stored: base(g, i) is
// recomputed from four hashes wherever it is needed.
//
// CTR synthesis follows what utree-build_gg + utree-compress do to such
// genomes: k-mers kept under the complevel rule (the `lv` bases before the
// 32-mer are A,G,C,T in that order, itree.c:605-616); a k-mer seen in several
// genomes gets the fold of the builder's relabelling rule applied in genome
// order (xeTreeU_RF, itree.c:285-305: cut at the last shared ';', give up below
// two shared ';'); records sorted by word; BinIx as XT_cmp32 builds it
// (itree.c:1281-1289, first-bin quirk included); label tail "label\tcount\n".
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define NUMBINS ((1u << 24) + 1u)

static thread_local char g_err[512];
#define SCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            snprintf(g_err, sizeof g_err, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e_),       \
                     __FILE__, __LINE__, cudaGetErrorString(e_));                                    \
            return 5;                                                                                \
        }                                                                                            \
    } while (0)

struct Universe {
    uint64_t seed;
    uint32_t n_phyla, n_genera, n_species, n_strains;   // per parent
    uint32_t genome_len;
};
__host__ __device__ inline uint32_t u_genomes(const Universe &u) { return u.n_phyla * u.n_genera * u.n_species * u.n_strains; }

__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline uint64_t h4(uint64_t seed, uint32_t level, uint32_t id, uint32_t i) {
    return mix64(seed ^ mix64(((uint64_t)level << 56) ^ ((uint64_t)id << 32) ^ i));
}
// base i of genome g (codes A=0 C=1 G=2 T=3)
__host__ __device__ inline uint32_t base_of(const Universe &u, uint32_t g, uint32_t i) {
    uint32_t sp = g / u.n_strains, ge = sp / u.n_species, ph = ge / u.n_genera;
    uint32_t b = (uint32_t)(h4(u.seed, 0, ph, i) & 3u);
    uint64_t x = h4(u.seed, 1, ge, i);
    if (x % 1000u < 100u) b = (b + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
    x = h4(u.seed, 2, sp, i);
    if (x % 1000u < 30u) b = (b + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
    x = h4(u.seed, 3, g, i);
    if (x % 1000u < 5u) b = (b + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
    return b;
}

// ---------------------------------------------------------------------------
// k-mer extraction (complevel rule) -> unordered (word, genome) pairs
// ---------------------------------------------------------------------------
#define CHUNK 512u
__global__ void __launch_bounds__(256)
extract_kernel(Universe u, uint32_t lv, uint32_t g0, uint32_t n_g, uint64_t *__restrict__ words, uint32_t *__restrict__ gids,
               unsigned long long *__restrict__ counter, uint64_t cap) {
    const uint32_t chunks = (u.genome_len + CHUNK - 1) / CHUNK;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < (uint64_t)n_g * chunks;
    uint32_t g = g0 + (uint32_t)(active ? t / chunks : 0), c = (uint32_t)(active ? t % chunks : 0);
    uint32_t lo = c * CHUNK, hi = min(lo + CHUNK, u.genome_len);        // window END positions [lo, hi)
    const uint32_t kv = 31u + lv;
    uint64_t w = 0;
    uint32_t hist = 0;
    const uint32_t hmask = lv ? ((1u << (2u * lv)) - 1u) : 0u;
    const uint32_t want = lv == 0 ? 0u : lv == 1 ? 0x0u : lv == 2 ? 0x2u : lv == 3 ? 0x9u : 0x27u;   // A,G,C,T
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k = 0; k < CHUNK + kv; ++k) {                        // uniform trip count across the warp
        int64_t i = (int64_t)lo - (int64_t)kv + (int64_t)k;            // kv bases of warm-up before the chunk
        bool emit = false;
        if (active && i >= 0 && i < (int64_t)hi) {
            uint32_t b = base_of(u, g, (uint32_t)i);
            hist = ((hist << 2) | (uint32_t)(w >> 62)) & hmask;
            w = (w << 2) | b;
            emit = i >= (int64_t)lo && i >= (int64_t)kv && hist == want;
        }
        uint32_t m = __ballot_sync(0xFFFFFFFFu, emit);
        if (m) {
            unsigned long long base = 0;
            int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (emit) {
                uint64_t idx = base + __popc(m & ((1u << lane) - 1u));
                if (idx < cap) { words[idx] = w; gids[idx] = g; }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// per run of equal words: fold the builder's relabelling rule in genome order
// ---------------------------------------------------------------------------
__device__ inline uint32_t common_ranks(const Universe &u, uint32_t a, uint32_t b) {
    uint32_t sa = a / u.n_strains, sb = b / u.n_strains;
    if (sa == sb) return 7;
    uint32_t ga = sa / u.n_species, gb = sb / u.n_species;
    if (ga == gb) return 6;
    if (ga / u.n_genera == gb / u.n_genera) return 5;
    return 1;
}
// label id of the depth-d prefix of genome g's taxonomy (d in 2..8)
__host__ __device__ inline uint32_t label_id(const Universe &u, uint32_t g, uint32_t d) {
    uint32_t G = u_genomes(u), S = G / u.n_strains, Q = S / u.n_species;
    uint32_t sp = g / u.n_strains, ge = sp / u.n_species, ph = ge / u.n_genera;
    if (d == 8) return g;
    if (d == 7) return G + sp;
    if (d == 6) return G + S + ge;
    return G + S + Q + ph * 4u + (5u - d);                           // f, o, c, p
}
__global__ void __launch_bounds__(256)
fold_kernel(Universe u, const uint64_t *__restrict__ words, const uint32_t *__restrict__ gids, uint64_t n,
            uint32_t *__restrict__ label, uint8_t *__restrict__ keep) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t w = words[i];
    if (i && words[i - 1] == w) { keep[i] = 0; label[i] = 0xFFFFFFFFu; return; }   // not a run head
    uint64_t e = i + 1;
    while (e < n && words[e] == w) ++e;
    // genomes of the run in ascending order, multiplicities included
    uint32_t g0 = 0xFFFFFFFFu;
    for (uint64_t j = i; j < e; ++j) g0 = min(g0, gids[j]);
    uint32_t d = 8, last = 0;
    bool bad = false, first = true;
    for (;;) {
        uint32_t cur = 0xFFFFFFFFu;
        if (first) cur = g0;
        else for (uint64_t j = i; j < e; ++j) { uint32_t x = gids[j]; if (x > last && x < cur) cur = x; }
        if (cur == 0xFFFFFFFFu) break;
        uint32_t mult = 0;
        for (uint64_t j = i; j < e; ++j) mult += gids[j] == cur;
        uint32_t steps = first ? 0u : mult;                           // repeats inside genome g0 are no-ops (same ix)
        for (uint32_t s = 0; s < steps && !bad; ++s) {
            uint32_t cr = common_ranks(u, g0, cur), m = min(d, cr);
            d = m < d ? m : d - 1;                                    // cut at the last shared ';'
            if (d < 2) bad = true;                                    // critical_cutoff (itree.c:295)
        }
        first = false; last = cur;
        if (bad) break;
    }
    keep[i] = !bad;
    label[i] = bad ? 0xFFFFFFFFu : label_id(u, g0, d);
}

__global__ void binix_kernel(const uint64_t *__restrict__ words, uint64_t n, uint32_t *__restrict__ binix) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = (uint32_t)(words[i] >> 40);
    if (i == 0) { for (uint32_t v = 0; v <= p; ++v) binix[v] = 0; }
    else {
        uint32_t q = (uint32_t)(words[i - 1] >> 40);
        for (uint32_t v = q + 1; v <= p; ++v) binix[v] = (uint32_t)i;
    }
    if (i == n - 1) for (uint32_t v = p + 1; v < NUMBINS; ++v) binix[v] = (uint32_t)n;
}
// XT_cmp32's first-bin quirk (itree.c:1284-1288): a first bin with ONE record is lost
__global__ void quirk_kernel(const uint64_t *__restrict__ words, uint64_t n, uint32_t *__restrict__ binix) {
    if (blockIdx.x || threadIdx.x || n < 2) return;
    uint32_t v0 = (uint32_t)(words[0] >> 40), v1 = (uint32_t)(words[1] >> 40);
    if (v1 == v0) return;                                             // >= 2 records in the first bin: fine
    for (uint32_t v = v0 + 1; v <= v1; ++v) binix[v] = 0;
}

__global__ void records_kernel(const uint64_t *__restrict__ words, const uint32_t *__restrict__ label, uint64_t n,
                               uint32_t ix_bytes, uint8_t *__restrict__ recs, unsigned long long *__restrict__ counts) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t w = words[i];
    uint32_t l = label[i];
    uint8_t *r = recs + i * (5 + ix_bytes);
    r[0] = (uint8_t)w; r[1] = (uint8_t)(w >> 8); r[2] = (uint8_t)(w >> 16); r[3] = (uint8_t)(w >> 24); r[4] = (uint8_t)(w >> 32);
    r[5] = (uint8_t)l; r[6] = (uint8_t)(l >> 8);
    if (ix_bytes == 4) { r[7] = (uint8_t)(l >> 16); r[8] = (uint8_t)(l >> 24); }
    atomicAdd(counts + l, 1ull);
}

static std::string tax_of(const Universe &u, uint32_t g, uint32_t depth) {
    uint32_t sp = g / u.n_strains, ge = sp / u.n_species, ph = ge / u.n_genera;
    char buf[256];
    const char *fmt[8] = {"k__Bacteria", ";p__P%u", ";c__C%u", ";o__O%u", ";f__F%u", ";g__G%u", ";s__S%u", ";t__T%u"};
    uint32_t val[8] = {0, ph, ph, ph, ph, ge, sp, g};
    std::string s;
    for (uint32_t d = 0; d < depth; ++d) { snprintf(buf, sizeof buf, fmt[d], val[d]); s += buf; }
    return s;
}

static int write_all(FILE *f, const void *p, size_t n) { return fwrite(p, 1, n, f) == n ? 0 : 2; }
static int dump_device(FILE *f, const void *d, size_t n) {
    const size_t CH = (size_t)128 << 20;
    void *pin;
    SCK(cudaMallocHost(&pin, CH));
    for (size_t o = 0; o < n; o += CH) {
        size_t c = n - o < CH ? n - o : CH;
        SCK(cudaMemcpy(pin, (const char *)d + o, c, cudaMemcpyDeviceToHost));
        if (write_all(f, pin, c)) { cudaFreeHost(pin); snprintf(g_err, sizeof g_err, "short write"); return 2; }
    }
    cudaFreeHost(pin);
    return 0;
}

extern "C" const char *uts_last_error(void) { return g_err; }

// Builds the CTR of the universe at `complevel` into out_path.  Returns 0 ok.
extern "C" int uts_build_ctr(int device, uint64_t seed, uint32_t n_phyla, uint32_t n_genera, uint32_t n_species,
                             uint32_t n_strains, uint32_t genome_len, uint32_t complevel, uint32_t ix_bytes,
                             const char *out_path, uint64_t *n_records, uint32_t *n_labels) {
    Universe u{seed, n_phyla, n_genera, n_species, n_strains, genome_len};
    if (complevel > 4 || (ix_bytes != 2 && ix_bytes != 4) || genome_len < 64) { snprintf(g_err, sizeof g_err, "bad argument"); return 1; }
    SCK(cudaSetDevice(device));
    const uint32_t G = u_genomes(u), S = G / n_strains, Q = S / n_species;
    const uint32_t L = G + S + Q + n_phyla * 4;
    if (ix_bytes == 2 && L > 65534) { snprintf(g_err, sizeof g_err, "%u labels do not fit uint16_t", L); return 1; }
    // expected k-mers: (len-31-lv)/4^lv per genome; 12 % head-room
    double expect = (double)G * (double)(genome_len - 31 - complevel) / (double)(1u << (2 * complevel));
    uint64_t cap = (uint64_t)(expect * 1.12) + (1u << 20);
    uint64_t *w_a, *w_b; uint32_t *g_a, *g_b; unsigned long long *counter;
    SCK(cudaMalloc(&w_a, cap * 8)); SCK(cudaMalloc(&w_b, cap * 8));
    SCK(cudaMalloc(&g_a, cap * 4)); SCK(cudaMalloc(&g_b, cap * 4));
    SCK(cudaMalloc(&counter, 8)); SCK(cudaMemset(counter, 0, 8));
    const uint32_t chunks = (genome_len + CHUNK - 1) / CHUNK;
    const uint32_t G_STEP = 256;
    for (uint32_t g0 = 0; g0 < G; g0 += G_STEP) {
        uint32_t ng = G - g0 < G_STEP ? G - g0 : G_STEP;
        uint64_t threads = (uint64_t)ng * chunks;
        extract_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(u, complevel, g0, ng, w_a, g_a, counter, cap);
    }
    SCK(cudaGetLastError());
    unsigned long long n_raw = 0;
    SCK(cudaMemcpy(&n_raw, counter, 8, cudaMemcpyDeviceToHost));
    if (n_raw > cap) { snprintf(g_err, sizeof g_err, "k-mer buffer overflow (%llu > %llu)", n_raw, (unsigned long long)cap); return 6; }
    // sort by word
    cub::DoubleBuffer<uint64_t> dk(w_a, w_b);
    cub::DoubleBuffer<uint32_t> dv(g_a, g_b);
    void *tmp = nullptr; size_t tmp_bytes = 0;
    SCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (uint64_t)n_raw));
    SCK(cudaMalloc(&tmp, tmp_bytes));
    SCK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (uint64_t)n_raw));
    SCK(cudaFree(tmp));
    uint64_t *ws = dk.Current(), *w_out = dk.Alternate();
    uint32_t *gs = dv.Current(), *l_out = dv.Alternate();
    uint32_t *lab; uint8_t *keep;
    SCK(cudaMalloc(&lab, n_raw * 4)); SCK(cudaMalloc(&keep, n_raw));
    fold_kernel<<<(unsigned)((n_raw + 255) / 256), 256>>>(u, ws, gs, n_raw, lab, keep);
    SCK(cudaGetLastError());
    unsigned long long *d_nsel;
    SCK(cudaMalloc(&d_nsel, 8));
    tmp = nullptr; tmp_bytes = 0;
    SCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, ws, keep, w_out, d_nsel, (uint64_t)n_raw));
    SCK(cudaMalloc(&tmp, tmp_bytes));
    SCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, ws, keep, w_out, d_nsel, (uint64_t)n_raw));
    SCK(cub::DeviceSelect::Flagged(tmp, tmp_bytes, lab, keep, l_out, d_nsel, (uint64_t)n_raw));
    SCK(cudaFree(tmp));
    unsigned long long n = 0;
    SCK(cudaMemcpy(&n, d_nsel, 8, cudaMemcpyDeviceToHost));
    if (!n || n >= 0xFFFFFFFFull) { snprintf(g_err, sizeof g_err, "unsupported record count %llu", n); return 6; }
    SCK(cudaFree(lab)); SCK(cudaFree(keep));
    uint32_t *binix; uint8_t *recs; unsigned long long *counts;
    const uint32_t sz = 5 + ix_bytes;
    SCK(cudaMalloc(&binix, (size_t)NUMBINS * 4));
    SCK(cudaMalloc(&recs, (size_t)n * sz));
    SCK(cudaMalloc(&counts, (size_t)L * 8)); SCK(cudaMemset(counts, 0, (size_t)L * 8));
    binix_kernel<<<(unsigned)((n + 255) / 256), 256>>>(w_out, n, binix);
    quirk_kernel<<<1, 32>>>(w_out, n, binix);
    records_kernel<<<(unsigned)((n + 255) / 256), 256>>>(w_out, l_out, n, ix_bytes, recs, counts);
    SCK(cudaGetLastError());
    SCK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h_counts(L);
    SCK(cudaMemcpy(h_counts.data(), counts, (size_t)L * 8, cudaMemcpyDeviceToHost));
    // file: header, BinIx, records, label tail (App. A)
    FILE *f = fopen(out_path, "wb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot create %s", out_path); return 2; }
    uint64_t md[4] = {8, 0, ix_bytes, n};
    int rc = write_all(f, md, 32);
    if (!rc) rc = dump_device(f, binix, (size_t)NUMBINS * 4);
    if (!rc) rc = dump_device(f, recs, (size_t)n * sz);
    if (!rc) {
        std::string tail;
        tail.reserve((size_t)L * 96);
        char num[32];
        for (uint32_t id = 0; id < L; ++id) {
            std::string s;
            if (id < G) s = tax_of(u, id, 8);
            else if (id < G + S) s = tax_of(u, (id - G) * n_strains, 7);
            else if (id < G + S + Q) s = tax_of(u, (id - G - S) * n_species * n_strains, 6);
            else { uint32_t k = id - G - S - Q; s = tax_of(u, (k / 4) * n_genera * n_species * n_strains, 5 - (k % 4)); }
            snprintf(num, sizeof num, "\t%llu\n", h_counts[id]);
            tail += s; tail += num;
        }
        rc = write_all(f, tail.data(), tail.size());
    }
    if (fclose(f) && !rc) rc = 2;
    if (rc && !g_err[0]) snprintf(g_err, sizeof g_err, "write error on %s", out_path);
    cudaFree(w_a); cudaFree(w_b); cudaFree(g_a); cudaFree(g_b); cudaFree(counter); cudaFree(d_nsel);
    cudaFree(binix); cudaFree(recs); cudaFree(counts);
    if (n_records) *n_records = n;
    if (n_labels) *n_labels = L;
    return rc;
}

// ---------------------------------------------------------------------------
// reads: fixed-width FASTA records ">r%09u\n" + bases + "\n"
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reads_kernel(Universe u, uint64_t rseed, uint64_t first, uint64_t n_reads, uint32_t read_len, uint32_t sub_permille,
             uint32_t n_permille, uint32_t random_permille, char *__restrict__ out) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t rec = 12u + read_len + 1u;                         // ">r" + 9 digits + "\n" + bases + "\n"
    char *p = out + r * rec;
    uint64_t id = first + r;
    p[0] = '>'; p[1] = 'r';
    uint64_t v = id;
    for (int k = 8; k >= 0; --k) { p[2 + k] = (char)('0' + v % 10); v /= 10; }
    p[11] = '\n';
    uint64_t h = mix64(rseed ^ mix64(id));
    uint32_t g = (uint32_t)(h % u_genomes(u));
    uint32_t start = (uint32_t)((h >> 32) % (u.genome_len - read_len + 1));
    bool rc = (mix64(h) >> 7) & 1u;
    bool is_random = (mix64(h ^ 0x55) % 1000u) < random_permille;
    bool has_n = (mix64(h ^ 0xAA) % 1000u) < n_permille;
    uint32_t n_pos = (uint32_t)(mix64(h ^ 0xAB) % read_len);
    const char A[4] = {'A', 'C', 'G', 'T'};
    for (uint32_t j = 0; j < read_len; ++j) {
        uint32_t src = rc ? read_len - 1 - j : j;                    // reverse-complemented reads walk backwards
        uint64_t x = mix64(h + 0x1000 + src);
        uint32_t b = is_random ? (uint32_t)(x >> 40) & 3u : base_of(u, g, start + src);
        if (!is_random && x % 1000u < sub_permille) b = (b + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
        if (rc) b = 3u - b;
        p[12 + j] = (has_n && j == n_pos) ? 'N' : A[b];
    }
    p[12 + read_len] = '\n';
}

extern "C" uint64_t uts_reads_bytes(uint64_t n_reads, uint32_t read_len) { return n_reads * (12ull + read_len + 1ull); }

// Fills host buffer `out` (uts_reads_bytes(n_reads, read_len) bytes) with reads [first, first+n_reads).
extern "C" int uts_make_reads(int device, uint64_t seed, uint32_t n_phyla, uint32_t n_genera, uint32_t n_species,
                              uint32_t n_strains, uint32_t genome_len, uint64_t read_seed, uint64_t first,
                              uint64_t n_reads, uint32_t read_len, uint32_t sub_permille, uint32_t n_permille,
                              uint32_t random_permille, char *out) {
    Universe u{seed, n_phyla, n_genera, n_species, n_strains, genome_len};
    if (!out || read_len < 1 || read_len > genome_len || first + n_reads > 999999999ull) { snprintf(g_err, sizeof g_err, "bad argument"); return 1; }
    SCK(cudaSetDevice(device));
    const uint64_t STEP = (uint64_t)1 << 20;
    const uint32_t rec = 12u + read_len + 1u;
    char *d;
    SCK(cudaMalloc(&d, STEP * rec));
    for (uint64_t o = 0; o < n_reads; o += STEP) {
        uint64_t c = n_reads - o < STEP ? n_reads - o : STEP;
        reads_kernel<<<(unsigned)((c + 255) / 256), 256>>>(u, read_seed, first + o, c, read_len, sub_permille, n_permille, random_permille, d);
        SCK(cudaGetLastError());
        SCK(cudaMemcpy(out + o * rec, d, c * rec, cudaMemcpyDeviceToHost));
    }
    cudaFree(d);
    return 0;
}

// ---------------------------------------------------------------------------
// long reads / whole-genome queries: variable-length records ">r%09u\n" + bases + "\n"
// ---------------------------------------------------------------------------
// One thread per output byte: the record of a byte is found by bisection over the record offsets.  A
// read no longer than a genome is a window of one genome; a longer one (whole-genome queries up to the
// 16,777,214-base limit) runs on through the following genomes.  Substitutions and strand as in reads_kernel.
__global__ void __launch_bounds__(256)
long_reads_kernel(Universe u, uint64_t rseed, uint64_t first, uint32_t n_reads, const uint64_t *__restrict__ off,
                  const uint32_t *__restrict__ len, uint64_t b0, uint64_t b1, uint32_t sub_permille, char *__restrict__ out) {
    const char A[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t i = b0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b1; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_reads;                                // off[lo] <= i < off[hi]
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
        const uint64_t local = i - off[lo], id = first + lo;
        const uint32_t L = len[lo];
        char c;
        if (local == 0) c = '>';
        else if (local == 1) c = 'r';
        else if (local < 11) { uint64_t v = id; for (uint64_t k = 10; k > local; --k) v /= 10; c = (char)('0' + v % 10); }
        else if (local == 11 || local == 12ull + L) c = '\n';
        else {
            const uint64_t j = local - 12, h = mix64(rseed ^ mix64(id));
            const uint32_t g0 = (uint32_t)(h % u_genomes(u));
            const uint64_t start = L <= u.genome_len ? (h >> 32) % (u.genome_len - L + 1) : 0;
            const bool rc = (mix64(h) >> 7) & 1u;
            const uint64_t src = rc ? L - 1 - j : j, p = start + src;
            const uint64_t x = mix64(h + 0x1000 + src);
            uint32_t b = base_of(u, (uint32_t)((g0 + p / u.genome_len) % u_genomes(u)), (uint32_t)(p % u.genome_len));
            if (x % 1000u < sub_permille) b = (b + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
            if (rc) b = 3u - b;
            c = A[b];
        }
        out[i - b0] = c;
    }
}
// Fills host buffer `out` (sum of 12 + len[r] + 1 bytes) with reads [first, first + n_reads) of the given lengths.
extern "C" int uts_make_long_reads(int device, uint64_t seed, uint32_t n_phyla, uint32_t n_genera, uint32_t n_species,
                                   uint32_t n_strains, uint32_t genome_len, uint64_t read_seed, uint64_t first,
                                   uint32_t n_reads, const uint32_t *len, uint32_t sub_permille, char *out) {
    Universe u{seed, n_phyla, n_genera, n_species, n_strains, genome_len};
    if (!out || !len || !n_reads || first + n_reads > 999999999ull) { snprintf(g_err, sizeof g_err, "bad argument"); return 1; }
    SCK(cudaSetDevice(device));
    std::vector<uint64_t> off(n_reads + 1);
    off[0] = 0;
    for (uint32_t r = 0; r < n_reads; ++r) {
        if (len[r] < 1 || len[r] > 16777214u) { snprintf(g_err, sizeof g_err, "read %u: bad length %u", r, len[r]); return 1; }
        off[r + 1] = off[r] + 12ull + len[r] + 1ull;
    }
    const uint64_t total = off[n_reads], STEP = (uint64_t)256 << 20;
    uint64_t *d_off; uint32_t *d_len; char *d;
    SCK(cudaMalloc(&d_off, (n_reads + 1) * 8ull)); SCK(cudaMalloc(&d_len, n_reads * 4ull)); SCK(cudaMalloc(&d, STEP));
    SCK(cudaMemcpy(d_off, off.data(), (n_reads + 1) * 8ull, cudaMemcpyHostToDevice));
    SCK(cudaMemcpy(d_len, len, n_reads * 4ull, cudaMemcpyHostToDevice));
    for (uint64_t o = 0; o < total; o += STEP) {
        const uint64_t c = total - o < STEP ? total - o : STEP;
        long_reads_kernel<<<148 * 16, 256>>>(u, read_seed, first, n_reads, d_off, d_len, o, o + c, sub_permille, d);
        SCK(cudaGetLastError());
        SCK(cudaMemcpy(out + o, d, c, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_off); cudaFree(d_len); cudaFree(d);
    return 0;
}

// ---------------------------------------------------------------------------
// a CTR with >= 2^32 - 1 records (8-byte BinIx entries, itree.c:757, 1303): closed-form content, so that lookups
// can be checked against arithmetic instead of a second 30 GB copy in the checker
// ---------------------------------------------------------------------------
// word(i) = i * S + mix64(i) % S with S = floor(2^64 / n): strictly increasing; label(i) = mix64(i ^ 0x5555) % n_labels.
__host__ __device__ inline uint64_t big_word(uint64_t i, uint64_t S) { return i * S + mix64(i) % S; }
__global__ void __launch_bounds__(256)
big_records_kernel(uint64_t i0, uint64_t count, uint64_t S, uint32_t n_labels, uint8_t *__restrict__ out) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const uint64_t i = i0 + k, w = big_word(i, S);
    const uint32_t lab = (uint32_t)(mix64(i ^ 0x5555ull) % n_labels);
    uint8_t *r = out + k * 7;
    for (int b = 0; b < 5; ++b) r[b] = (uint8_t)(w >> (8 * b));
    r[5] = (uint8_t)lab; r[6] = (uint8_t)(lab >> 8);
}
__global__ void __launch_bounds__(256)
big_binix_kernel(uint64_t n, uint64_t S, uint64_t *__restrict__ binix) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > (1ull << 24)) return;
    if (p == (1ull << 24)) { binix[p] = n; return; }
    const uint64_t T = p << 40, c = T / S;                          // first record whose word is >= T
    uint64_t f = c < n && big_word(c, S) >= T ? c : c + 1;
    binix[p] = f < n ? f : n;
}
extern "C" int uts_build_big_ctr(int device, uint64_t n, uint32_t n_labels, const char *out_path) {
    if (!out_path || n < 2 || !n_labels || n_labels > 65000) { snprintf(g_err, sizeof g_err, "bad argument"); return 1; }
    SCK(cudaSetDevice(device));
    const uint64_t S = ~0ull / n;
    FILE *f = fopen(out_path, "wb");
    if (!f) { snprintf(g_err, sizeof g_err, "cannot open %s", out_path); return 2; }
    int rc = 0;
    const uint64_t md[4] = {8, 0, 2, n};
    rc |= fwrite(md, 8, 4, f) != 4;
    {
        uint64_t *d_b;
        SCK(cudaMalloc(&d_b, ((1ull << 24) + 1) * 8));
        big_binix_kernel<<<(unsigned)(((1ull << 24) + 1 + 255) / 256), 256>>>(n, S, d_b);
        SCK(cudaGetLastError());
        std::vector<uint64_t> h((1ull << 24) + 1);
        SCK(cudaMemcpy(h.data(), d_b, h.size() * 8, cudaMemcpyDeviceToHost));
        cudaFree(d_b);
        if (n < 0xFFFFFFFFull) { std::vector<uint32_t> h32(h.begin(), h.end()); rc |= fwrite(h32.data(), 4, h32.size(), f) != h32.size(); }
        else rc |= fwrite(h.data(), 8, h.size(), f) != h.size();
    }
    {
        const uint64_t STEP = (uint64_t)32 << 20;                   // records per chunk
        uint8_t *d; void *hbuf;
        SCK(cudaMalloc(&d, STEP * 7)); SCK(cudaMallocHost(&hbuf, STEP * 7));
        for (uint64_t i0 = 0; i0 < n && !rc; i0 += STEP) {
            const uint64_t c = n - i0 < STEP ? n - i0 : STEP;
            big_records_kernel<<<(unsigned)((c + 255) / 256), 256>>>(i0, c, S, n_labels, d);
            SCK(cudaGetLastError());
            SCK(cudaMemcpy(hbuf, d, c * 7, cudaMemcpyDeviceToHost));
            rc |= fwrite(hbuf, 7, c, f) != c;
        }
        cudaFree(d); cudaFreeHost(hbuf);
    }
    for (uint32_t l = 0; l < n_labels && !rc; ++l) rc |= fprintf(f, "k__K;p__P%u;c__C%u\t1\n", l % 97, l) < 0;
    if (fclose(f)) rc = 1;
    if (rc) snprintf(g_err, sizeof g_err, "write error on %s", out_path);
    return rc ? 2 : 0;
}

// One genome as ASCII (for building toy trees with the reference builder and
// for cross-checking base_of on the host).
extern "C" void uts_genome_ascii(uint64_t seed, uint32_t n_phyla, uint32_t n_genera, uint32_t n_species,
                                 uint32_t n_strains, uint32_t genome_len, uint32_t g, char *out) {
    Universe u{seed, n_phyla, n_genera, n_species, n_strains, genome_len};
    for (uint32_t i = 0; i < genome_len; ++i) out[i] = "ACGT"[base_of(u, g, i)];
}
extern "C" void uts_genome_tax(uint64_t seed, uint32_t n_phyla, uint32_t n_genera, uint32_t n_species,
                               uint32_t n_strains, uint32_t genome_len, uint32_t g, char *out, size_t cap) {
    Universe u{seed, n_phyla, n_genera, n_species, n_strains, genome_len};
    std::string s = tax_of(u, g, 8);
    snprintf(out, cap, "%s", s.c_str());
}
