// zb_deflate.cu -- K1..K3: the DEFLATE compressor as a pipeline of data-parallel kernels.
//
// Replaces, in the reference: fill_window/INSERT_STRING/longest_match/deflate_fast/deflate_slow
// (qcsrc/deflate.c:1266,189,1027,1448,1554) and _tr_tally/_tr_flush_block/build_tree/gen_bitlen/
// gen_codes/send_all_trees/compress_block/send_bits (qcsrc/trees.c:1022,921,619,490,577,838,1072,217).
//
// The reference interleaves these per input byte inside one sequential loop.  Here the same
// decisions are taken in five passes over HBM-resident arrays, each with its own parallel axis:
//
//   K1a link   one warp per 128 KiB segment walks it in order with a 15-bit hash -> last position
//              table in shared memory (the reference's head[]), 32 positions per step, and writes for
//              every position the distance to the previous position with the same 3-byte hash (the
//              reference's prev[] chain, stored as deltas so chains cross chunk boundaries and the
//              32 KiB of history in front of a chunk needs no copy -- it is simply there).
//   K1b match  one thread per input position follows that chain up to max_chain candidates
//              (configuration_table, deflate.c:137-149) and keeps the longest match, with the
//              reference's quick reject on the byte that would extend the best match so far.
//   K1c parse  one warp per chunk turns per-position matches into the token stream: greedy
//              (deflate_fast) or lazy with max_lazy and TOO_FAR (deflate_slow, deflate.c:1601-1612);
//              the serial "next position" recurrence is resolved 32 positions at a time with shuffles.
//              Tokens are tallied into per-block histograms (shared-memory atomics), blocks close every
//              16 Ki symbols like the reference's lit_bufsize.
//   K2 codes   one warp per block builds the three length-limited canonical codes exactly as trees.c
//              does (same heap order, same tie-break, same overflow repair -- unit-tested against the
//              oracle), prices stored / fixed / dynamic with the reference's rule (trees.c:955-1001) and
//              serialises the dynamic header.
//   plan+scan  exact compressed size of every chunk -> exclusive prefix sum -> byte offsets.
//   K3 pack    one CTA per chunk: per-symbol (code, length) -> prefix sum of lengths -> bits OR-ed into a
//              shared-memory staging window -> bytes to their final place.  Chunks end byte-aligned (empty
//              stored block, i.e. what Z_SYNC_FLUSH emits, deflate.c:808-819), so shards from several
//              GPUs concatenate by byte copy.
//
// Output is a valid DEFLATE stream, not the reference's bytes: parity is "reference inflate decodes
// it bit-exact" plus a size bound (<= 1.02 x reference at the same level), see tests/.
#include "zb_deflate.cuh"
#include "zb200_internal.h"

namespace zb {

static const LevelCfg h_levels[10] = {
    {0, 0, 0, 0, 0},      {4, 4, 8, 4, 1},       {4, 5, 16, 8, 1},     {4, 6, 32, 32, 1},
    {4, 4, 16, 16, 2},    {8, 16, 32, 32, 2},    {8, 16, 128, 128, 2}, {8, 32, 128, 256, 2},
    {32, 128, 258, 1024, 2}, {32, 258, 258, 4096, 2}};

constexpr uint32_t kFullMask = 0xffffffffu;
constexpr int kHashBits = 15;

// ------------------------------------------------------------------------------------------
// K1a: hash-chain links
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash3(uint32_t v) { return ((v & 0xffffffu) * 0x9E3779B1u) >> (32 - kHashBits); }

__global__ void __launch_bounds__(32)
k_lz_link(const uint8_t* __restrict__ buf, uint64_t total, uint16_t* __restrict__ dist16)
{
    extern __shared__ uint16_t s_head[];                       // 2^15 entries: low 16 bits of the last position
    const int lane = threadIdx.x;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i = lane; i < (1 << kHashBits) / 2; i += 32) reinterpret_cast<uint32_t*>(s_head)[i] = 0;
    __syncwarp();

    const uint64_t seg_beg = (uint64_t)blockIdx.x * kChunk;
    const uint64_t seg_end = min(total, seg_beg + kChunk);
    const uint64_t prime_beg = seg_beg > kWindow ? seg_beg - kWindow : 0;
    const uint32_t mis = (uint32_t)((uintptr_t)buf & 3);
    const uint32_t* words = reinterpret_cast<const uint32_t*>(buf - mis);
    const uint64_t nwords = (mis + total + 3) >> 2;

    // q = p + mis indexes bytes from the aligned base; tiles of 128 bytes, one word per lane.
    // Only the head[] read -> write chain is serial; loads are issued two tiles ahead and the hashes
    // and intra-step groups of a whole tile are computed before that chain starts.
    const uint64_t tile_beg = (prime_beg + mis) >> 7, tile_end = (seg_end + mis + 127) >> 7;
    const uint64_t q_lo = prime_beg + mis;                      // first position to insert
    const uint64_t q_hi = min(seg_end, total >= 2 ? total - 2 : 0) + mis;   // one past the last hashable position
    const uint64_t q_seg = seg_beg + mis, q_end = seg_end + mis;            // positions whose link is stored
    auto ldw = [&](uint64_t t) { const uint64_t w = t * 32 + lane; return w < nwords ? __ldg(words + w) : 0u; };
    uint32_t cur = ldw(tile_beg), nxt = ldw(tile_beg + 1);
    for (uint64_t tile = tile_beg; tile < tile_end; ++tile) {
        const uint32_t nxt2 = ldw(tile + 2);
        const uint32_t nxt0 = __shfl_sync(kFullMask, nxt, 0);
        const uint64_t qbase = tile * 128;
        // tile-relative bounds (clamped to [0,128])
        const int lo = (int)(q_lo > qbase ? min((uint64_t)128, q_lo - qbase) : 0);
        const int hi = (int)(q_hi > qbase ? min((uint64_t)128, q_hi - qbase) : 0);
        const int slo = (int)(q_seg > qbase ? min((uint64_t)128, q_seg - qbase) : 0);
        const int shi = (int)(q_end > qbase ? min((uint64_t)128, q_end - qbase) : 0);
        uint32_t h[4], d[4];
        bool valid[4], leader[4], from_head[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int li = 32 * k + lane, j = li >> 2;
            const uint32_t wa = __shfl_sync(kFullMask, cur, j);
            uint32_t wb = __shfl_sync(kFullMask, cur, (j + 1) & 31);
            if (j == 31) wb = nxt0;
            const uint32_t v = __funnelshift_r(wa, wb, (li & 3) * 8);
            valid[k] = li >= lo && li < hi;
            h[k] = valid[k] ? hash3(v) : (0x10000u + lane);
            const uint32_t grp = __match_any_sync(kFullMask, h[k]);
            const uint32_t lower = grp & lt;
            leader[k] = valid[k] && (grp >> lane) == 1u;         // highest lane of its group writes head[]
            from_head[k] = valid[k] && lower == 0;
            d[k] = (valid[k] && lower) ? (uint32_t)(lane - (31 - __clz(lower))) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t q32 = (uint32_t)qbase + 32 * k + lane;
            if (from_head[k]) {
                uint32_t dd = (q32 - s_head[h[k]]) & 0xffffu;
                if (dd == 0) dd = 0x10000u;
                const uint64_t p = qbase + 32 * k + lane - mis;
                d[k] = (dd > kWindow || dd > p) ? 0u : dd;
            }
            __syncwarp();
            if (leader[k]) s_head[h[k]] = (uint16_t)q32;
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int li = 32 * k + lane;
            if (li >= slo && li < shi) dist16[qbase + li - mis] = (uint16_t)d[k];
        }
        cur = nxt; nxt = nxt2;
    }
}

// ------------------------------------------------------------------------------------------
// K1b: longest match per position
// ------------------------------------------------------------------------------------------
struct WordView {
    const uint32_t* words; uint32_t mis; uint64_t last;         // last valid word index
    __device__ __forceinline__ uint32_t ld32(uint64_t p) const  // 4 bytes at byte position p (little endian)
    {
        const uint64_t q = p + mis, w = q >> 2;
        const uint32_t a = __ldg(words + w), b = __ldg(words + min(w + 1, last));
        return __funnelshift_r(a, b, (uint32_t)(q & 3) * 8);
    }
};

__global__ void __launch_bounds__(256)
k_lz_match(const uint8_t* __restrict__ buf, uint64_t total, uint64_t dict, const uint16_t* __restrict__ dist16,
           uint32_t* __restrict__ mt, int max_chain, int nice)
{
    const uint64_t rel = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const uint64_t p = dict + rel;
    if (p >= total) return;
    const uint64_t chunk_end = min(total, dict + (rel / kChunk + 1) * kChunk);
    const uint32_t maxlen = (uint32_t)min((uint64_t)kMaxMatch, chunk_end - p);
    if (maxlen < kMinMatch) { mt[rel] = 0; return; }
    WordView wv;
    wv.mis = (uint32_t)((uintptr_t)buf & 3);
    wv.words = reinterpret_cast<const uint32_t*>(buf - wv.mis);
    wv.last = ((wv.mis + total + 3) >> 2) - 1;
    const uint32_t nice_eff = min((uint32_t)nice, maxlen);

    uint32_t best_len = kMinMatch - 1, best_dist = 0, acc = 0;
    uint32_t d = dist16[p];
    const uint32_t head4 = wv.ld32(p);
    int chain = max_chain;
    while (d != 0 && chain-- > 0) {
        acc += d;
        if (acc > kWindow) break;
        const uint64_t cand = p - acc;
        // quick rejects: the byte that must match to beat best_len, then the first three (deflate.c:1121-1124)
        if (buf[cand + best_len] == buf[p + best_len] && ((wv.ld32(cand) ^ head4) & 0xffffffu) == 0) {
            uint32_t len = 3;
            while (len < maxlen) {
                const uint32_t x = wv.ld32(p + len) ^ wv.ld32(cand + len);
                if (x) { len += (__ffs(x) - 1) >> 3; break; }
                len += 4;
            }
            if (len > maxlen) len = maxlen;
            if (len > best_len) {
                best_len = len; best_dist = acc;
                if (len >= nice_eff) break;
            }
        }
        d = dist16[cand];
    }
    mt[rel] = best_len >= kMinMatch ? ((best_dist << 16) | best_len) : 0u;
}

// ------------------------------------------------------------------------------------------
// K1c: parse (greedy / lazy), tokens, histograms, block boundaries
// ------------------------------------------------------------------------------------------
constexpr int kParseWarps = 4;
constexpr int kHistSize = 320;                                   // 0..287 literal/length, 288..319 distance

__global__ void __launch_bounds__(kParseWarps * 32)
k_lz_parse(const uint8_t* __restrict__ src, uint64_t n, const uint32_t* __restrict__ mt, uint32_t* __restrict__ tok,
           BlockMeta* __restrict__ blk, uint32_t* __restrict__ blk_hist, ChunkMeta* __restrict__ chunks,
           int kind, uint32_t max_lazy)
{
    __shared__ uint32_t s_hist[kParseWarps][kHistSize];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t nchunks = (n + kChunk - 1) / kChunk;
    const uint64_t c = (uint64_t)blockIdx.x * kParseWarps + warp;
    if (c >= nchunks) return;
    uint32_t* hist = s_hist[warp];
    for (int i = lane; i < kHistSize; i += 32) hist[i] = 0;
    __syncwarp();

    const uint64_t cbeg = c * kChunk;
    const uint32_t clen = (uint32_t)min((uint64_t)kChunk, n - cbeg);
    const uint32_t* m = mt + cbeg;
    const uint8_t* in = src + cbeg;
    uint32_t* out = tok + cbeg;
    uint32_t pos = 0, ntok = 0, nblk = 0, blk_tok0 = 0, blk_in0 = 0;

    auto close_block = [&](uint32_t end_pos) {
        __syncwarp();
        uint32_t* g = blk_hist + (c * kMaxBlocks + nblk) * kHistSize;
        for (int i = lane; i < kHistSize; i += 32) { g[i] = hist[i] + (i == 256 ? 1u : 0u); hist[i] = 0; }
        if (lane == 0) blk[c * kMaxBlocks + nblk] = BlockMeta{blk_tok0, ntok - blk_tok0, blk_in0, end_pos - blk_in0, 0, 0, 0, 0};
        nblk++; blk_tok0 = ntok; blk_in0 = end_pos;
        __syncwarp();
    };

    while (pos < clen) {
        const uint32_t i = pos + lane;
        uint32_t mv = i < clen ? m[i] : 0u;
        uint32_t mn = __shfl_down_sync(kFullMask, mv, 1);
        const uint32_t m32 = (pos + 32 < clen) ? m[pos + 32] : 0u;     // uniform load
        if (lane == 31) mn = m32;
        uint32_t L = mv & 0x1ffu, dist = mv >> 16;
        bool take;
        if (kind == 2) {
            uint32_t Ln = mn & 0x1ffu;
            if (L == kMinMatch && dist > kTooFar) L = 0;
            if (Ln == kMinMatch && (mn >> 16) > kTooFar) Ln = 0;
            take = L >= kMinMatch && !(L < max_lazy && Ln > L);
        } else {
            take = L >= kMinMatch;
        }
        const uint32_t step = take ? L : 1u;
        // follow the "next position" recurrence through the window
        uint32_t cur = 0, mask = 0;
        while (cur < 32 && pos + cur < clen) {
            mask |= 1u << cur;
            cur += __shfl_sync(kFullMask, step, cur);
        }
        if (mask & (1u << lane)) {
            const uint32_t idx = ntok + __popc(mask & lt);
            if (take) {
                out[idx] = (dist << 16) | (L - kMinMatch);
                atomicAdd(&hist[257 + len_code(L - kMinMatch)], 1u);
                atomicAdd(&hist[288 + dist_code(dist - 1)], 1u);
            } else {
                const uint32_t b = in[i];
                out[idx] = b;
                atomicAdd(&hist[b], 1u);
            }
        }
        ntok += __popc(mask);
        pos += cur;
        if (ntok - blk_tok0 + 32 > kBlockTokens && pos < clen) close_block(pos);
    }
    close_block(clen);
    if (lane == 0) chunks[c] = ChunkMeta{nblk, 0, 0, 0};
}

// ------------------------------------------------------------------------------------------
// K2: length-limited canonical codes per block (trees.c:490-860), block pricing (trees.c:921-1001)
// ------------------------------------------------------------------------------------------
constexpr int kCodeWarps = 4;
constexpr int kTreeMax = 2 * 286 + 1;

__constant__ uint8_t c_bl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct TreeScratch {
    uint16_t freq[kTreeMax], up[kTreeMax], len[kTreeMax];
    uint16_t heap[kTreeMax + 1];
    uint8_t  depth[kTreeMax];
    uint16_t bl_count[16];
    uint16_t llen[288 + 2], dlen[32 + 2];                       // +sentinel slot for the run-length scan
    uint16_t blfreq[20], bllen[20], blcode[20];
    uint16_t lcode[288], dcode[32];
    int heap_len, heap_max;
};

struct BitSink {                                                // serial bit writer into global words
    uint32_t* w; uint64_t acc; int n; uint32_t count, total;
    __device__ void put(uint32_t v, int bits)
    {
        acc |= (uint64_t)v << n; n += bits; total += bits;
        if (n >= 32) { w[count++] = (uint32_t)acc; acc >>= 32; n -= 32; }
    }
    __device__ void finish() { if (n > 0) w[count++] = (uint32_t)acc; }
};

__device__ __forceinline__ bool node_lighter(const TreeScratch* t, int a, int b)
{
    return t->freq[a] < t->freq[b] || (t->freq[a] == t->freq[b] && t->depth[a] <= t->depth[b]);
}
__device__ void sift_down(TreeScratch* t, int k)
{
    const int v = t->heap[k];
    for (int j = k << 1; j <= t->heap_len; j <<= 1) {
        if (j < t->heap_len && node_lighter(t, t->heap[j + 1], t->heap[j])) j++;
        if (node_lighter(t, v, t->heap[j])) break;
        t->heap[k] = t->heap[j]; k = j;
    }
    t->heap[k] = (uint16_t)v;
}
__device__ __forceinline__ uint32_t bit_reverse(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// Serial (one lane).  freq[0..nsym) in t->freq; writes lengths to out_len[0..nsym) and codes to out_code.
// Returns max_code.
__device__ int make_code(TreeScratch* t, int nsym, int max_length, uint16_t* out_len, uint16_t* out_code)
{
    int max_code = -1;
    t->heap_len = 0; t->heap_max = kTreeMax;
    for (int n = 0; n < nsym; n++) {
        if (t->freq[n]) { t->heap[++t->heap_len] = (uint16_t)(max_code = n); t->depth[n] = 0; }
        else t->len[n] = 0;
    }
    while (t->heap_len < 2) {                                   // force two codes (trees.c:650-656)
        const int n = (max_code < 2 ? ++max_code : 0);
        t->heap[++t->heap_len] = (uint16_t)n;
        t->freq[n] = 1; t->depth[n] = 0;
    }
    for (int n = t->heap_len / 2; n >= 1; n--) sift_down(t, n);
    int node = nsym;
    do {
        const int n = t->heap[1];
        t->heap[1] = t->heap[t->heap_len--]; sift_down(t, 1);
        const int m = t->heap[1];
        t->heap[--t->heap_max] = (uint16_t)n; t->heap[--t->heap_max] = (uint16_t)m;
        t->freq[node] = (uint16_t)(t->freq[n] + t->freq[m]);
        t->depth[node] = (uint8_t)((t->depth[n] >= t->depth[m] ? t->depth[n] : t->depth[m]) + 1);
        t->up[n] = t->up[m] = (uint16_t)node;
        t->heap[1] = (uint16_t)node++; sift_down(t, 1);
    } while (t->heap_len >= 2);
    t->heap[--t->heap_max] = t->heap[1];

    // lengths with the reference's overflow repair (trees.c:490-567)
    for (int b = 0; b <= 15; b++) t->bl_count[b] = 0;
    int over = 0, h;
    t->len[t->heap[t->heap_max]] = 0;
    for (h = t->heap_max + 1; h < kTreeMax; h++) {
        const int n = t->heap[h];
        int bits = t->len[t->up[n]] + 1;
        if (bits > max_length) { bits = max_length; over++; }
        t->len[n] = (uint16_t)bits;
        if (n > max_code) continue;
        t->bl_count[bits]++;
    }
    if (over) {
        do {
            int bits = max_length - 1;
            while (t->bl_count[bits] == 0) bits--;
            t->bl_count[bits]--; t->bl_count[bits + 1] += 2; t->bl_count[max_length]--;
            over -= 2;
        } while (over > 0);
        for (int bits = max_length; bits != 0; bits--) {
            int n = t->bl_count[bits];
            while (n != 0) {
                const int m = t->heap[--h];
                if (m > max_code) continue;
                t->len[m] = (uint16_t)bits;
                n--;
            }
        }
    }
    // canonical codes (trees.c:577-609)
    uint32_t next[16], code = 0;
    for (int b = 1; b <= 15; b++) { code = (code + t->bl_count[b - 1]) << 1; next[b] = code; }
    for (int n = 0; n < nsym; n++) {
        const int l = n <= max_code ? t->len[n] : 0;
        out_len[n] = (uint16_t)l;
        out_code[n] = l ? (uint16_t)bit_reverse(next[l]++, l) : 0;
    }
    return max_code;
}

// Run-length walk over code lengths (trees.c:707-797): sink == nullptr tallies blfreq, else emits.
__device__ void walk_lengths(TreeScratch* t, uint16_t* lens, int max_code, BitSink* sink)
{
    int prevlen = -1, nextlen = lens[0], count = 0, maxc = 7, minc = 4;
    if (nextlen == 0) { maxc = 138; minc = 3; }
    lens[max_code + 1] = 0xffff;
    for (int n = 0; n <= max_code; n++) {
        const int cur = nextlen; nextlen = lens[n + 1];
        if (++count < maxc && cur == nextlen) continue;
        if (count < minc) {
            if (sink) { do sink->put(t->blcode[cur], t->bllen[cur]); while (--count); }
            else t->blfreq[cur] = (uint16_t)(t->blfreq[cur] + count);
        } else if (cur != 0) {
            if (cur != prevlen) {
                if (sink) { sink->put(t->blcode[cur], t->bllen[cur]); count--; }
                else t->blfreq[cur]++;
            }
            if (sink) { sink->put(t->blcode[16], t->bllen[16]); sink->put((uint32_t)count - 3, 2); }
            else t->blfreq[16]++;
        } else if (count <= 10) {
            if (sink) { sink->put(t->blcode[17], t->bllen[17]); sink->put((uint32_t)count - 3, 3); }
            else t->blfreq[17]++;
        } else {
            if (sink) { sink->put(t->blcode[18], t->bllen[18]); sink->put((uint32_t)count - 11, 7); }
            else t->blfreq[18]++;
        }
        count = 0; prevlen = cur;
        if (nextlen == 0) { maxc = 138; minc = 3; }
        else if (cur == nextlen) { maxc = 6; minc = 3; }
        else { maxc = 7; minc = 4; }
    }
}

__device__ __forceinline__ uint32_t fixed_lit_len(uint32_t s) { return s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8; }
__device__ __forceinline__ uint32_t fixed_lit_code(uint32_t s)
{
    const uint32_t c = s < 144 ? 0x30 + s : s < 256 ? 0x190 + (s - 144) : s < 280 ? (s - 256) : 0xC0 + (s - 280);
    return bit_reverse(c, (int)fixed_lit_len(s));
}

__global__ void __launch_bounds__(kCodeWarps * 32)
k_huff_build(uint64_t nchunks, const ChunkMeta* __restrict__ chunks, BlockMeta* __restrict__ blk,
             const uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ blk_codes, uint32_t* __restrict__ blk_hdr,
             int force_fixed)
{
    __shared__ TreeScratch s_t[kCodeWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t slot = (uint64_t)blockIdx.x * kCodeWarps + warp;
    const uint64_t c = slot / kMaxBlocks;
    if (c >= nchunks || (slot % kMaxBlocks) >= chunks[c].nblocks) return;
    TreeScratch* t = &s_t[warp];
    const uint32_t* hist = blk_hist + slot * kHistSize;
    uint32_t* codes = blk_codes + slot * kHistSize;

    // ---- literal/length and distance codes ----
    for (int i = lane; i < 286; i += 32) t->freq[i] = (uint16_t)hist[i];
    __syncwarp();
    int lmax = 0, dmax = 0;
    if (lane == 0) lmax = make_code(t, 286, 15, t->llen, t->lcode);
    __syncwarp();
    if (lane < 30) t->freq[lane] = (uint16_t)hist[288 + lane];
    __syncwarp();
    if (lane == 0) dmax = make_code(t, 30, 15, t->dlen, t->dcode);
    __syncwarp();
    lmax = __shfl_sync(kFullMask, lmax, 0); dmax = __shfl_sync(kFullMask, dmax, 0);

    // ---- cost of the data under the dynamic and the fixed code (all lanes) ----
    uint32_t dyn = 0, fix = 0;
    for (int s = lane; s < 286; s += 32) {
        const uint32_t f = hist[s];
        const uint32_t x = s >= 257 ? len_extra_bits((uint32_t)s - 257) : 0;
        dyn += f * (t->llen[s] + x); fix += f * (fixed_lit_len((uint32_t)s) + x);
    }
    if (lane < 30) {
        const uint32_t f = hist[288 + lane], x = dist_extra_bits((uint32_t)lane);
        dyn += f * (t->dlen[lane] + x); fix += f * (5 + x);
    }
#pragma unroll
    for (int k = 16; k; k >>= 1) { dyn += __shfl_xor_sync(kFullMask, dyn, k); fix += __shfl_xor_sync(kFullMask, fix, k); }

    // ---- code-length code and header (one lane) ----
    BlockMeta bm = blk[slot];
    uint32_t type = 0, body = 0, hdr_bits = 0;
    if (lane == 0) {
        for (int i = 0; i < 19; i++) t->blfreq[i] = 0;
        walk_lengths(t, t->llen, lmax, nullptr);
        walk_lengths(t, t->dlen, dmax, nullptr);
        for (int i = 0; i < 19; i++) t->freq[i] = t->blfreq[i];
        make_code(t, 19, 7, t->bllen, t->blcode);
        int max_bl = 18;
        while (max_bl >= 3 && t->bllen[c_bl_order[max_bl]] == 0) max_bl--;
        uint32_t tree_bits = 14 + 3 * (max_bl + 1);
        for (int i = 0; i < 19; i++) tree_bits += t->blfreq[i] * (t->bllen[i] + (i == 16 ? 2 : i == 17 ? 3 : i == 18 ? 7 : 0));
        const uint32_t opt_len = dyn + tree_bits, static_len = fix;
        uint32_t opt_lenb = (opt_len + 3 + 7) >> 3;
        const uint32_t static_lenb = (static_len + 3 + 7) >> 3;
        if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
        if (bm.in_len + 4 <= opt_lenb && !force_fixed) { type = 0; body = 0; }
        else if (static_lenb == opt_lenb || force_fixed) { type = 1; body = static_len; }
        else {
            type = 2; body = opt_len; hdr_bits = tree_bits;
            BitSink sink{blk_hdr + slot * kHdrWords, 0, 0, 0, 0};
            sink.put((uint32_t)lmax + 1 - 257, 5); sink.put((uint32_t)dmax + 1 - 1, 5); sink.put((uint32_t)max_bl + 1 - 4, 4);
            for (int r = 0; r <= max_bl; r++) sink.put(t->bllen[c_bl_order[r]], 3);
            walk_lengths(t, t->llen, lmax, &sink);
            walk_lengths(t, t->dlen, dmax, &sink);
            sink.finish();
            hdr_bits = sink.total;
        }
        t->llen[lmax + 1] = 0; t->dlen[dmax + 1] = 0;              // drop the run-length sentinels
        bm.type = type; bm.body_bits = body; bm.hdr_bits = hdr_bits;
        blk[slot] = bm;
    }
    type = __shfl_sync(kFullMask, type, 0);
    __syncwarp();
    // ---- code tables for the packer: code | len << 16 ----
    for (int s = lane; s < 288; s += 32) {
        uint32_t v;
        if (type == 1) v = fixed_lit_code((uint32_t)s) | (fixed_lit_len((uint32_t)s) << 16);
        else v = s < 286 ? (t->lcode[s] | ((uint32_t)(t->llen[s] & 0xff) << 16)) : 0;
        codes[s] = v;
    }
    {
        uint32_t v;
        if (type == 1) v = bit_reverse((uint32_t)lane, 5) | (5u << 16);
        else v = lane < 30 ? (t->dcode[lane] | ((uint32_t)(t->dlen[lane] & 0xff) << 16)) : 0;
        codes[288 + lane] = v;
    }
}

// ------------------------------------------------------------------------------------------
// plan + scan: exact chunk sizes and offsets
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t stored_chunk_bytes(uint32_t len) { return (uint64_t)len + 5ull * ((len + 65534u) / 65535u); }

__global__ void k_plan(uint64_t nchunks, uint64_t n, ChunkMeta* __restrict__ chunks, const BlockMeta* __restrict__ blk,
                       int last_is_final, int force_stored, int force_mark)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    ChunkMeta cm = chunks[c];
    const uint32_t clen = (uint32_t)min((uint64_t)kChunk, n - c * kChunk);
    const bool final_chunk = last_is_final && c == nchunks - 1;
    const bool mark = force_mark && !last_is_final && c == nchunks - 1;   // Z_SYNC_FLUSH marker wanted regardless
    uint64_t bytes;
    if (force_stored) {
        cm.stored = 1; bytes = stored_chunk_bytes(clen) + (mark ? 5 : 0);
    } else {
        uint64_t bits = 0;
        bool ends_stored = false;
        for (uint32_t b = 0; b < cm.nblocks; b++) {
            const BlockMeta bm = blk[c * kMaxBlocks + b];
            if (bm.type == 0) { bits = ((bits + 3 + 7) & ~7ull) + 32 + 8ull * bm.in_len; ends_stored = true; }
            else { bits += 3 + bm.body_bits; ends_stored = false; }
        }
        if (final_chunk || (ends_stored && !mark)) bytes = (bits + 7) >> 3;
        else bytes = ((bits + 3 + 7) >> 3) + 4;                 // empty stored block: 000, pad, 00 00 FF FF
        const uint64_t sb = stored_chunk_bytes(clen) + (mark ? 5 : 0);
        cm.stored = sb <= bytes ? 1u : 0u;
        if (cm.stored) bytes = sb;
    }
    cm.bytes = bytes;
    chunks[c] = cm;
}

// single CTA: exclusive prefix sum of chunk sizes; result[0] = total payload bytes
__global__ void __launch_bounds__(1024) k_scan(uint64_t nchunks, ChunkMeta* __restrict__ chunks, uint64_t base,
                                               uint64_t* __restrict__ result)
{
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = base;
    __syncthreads();
    for (uint64_t i0 = 0; i0 < nchunks; i0 += 1024) {
        const uint64_t i = i0 + threadIdx.x;
        const uint64_t v = i < nchunks ? chunks[i].bytes : 0;
        uint64_t x = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) { const uint64_t y = __shfl_up_sync(kFullMask, x, k); if (lane >= k) x += y; }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_w[lane];
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) { const uint64_t y = __shfl_up_sync(kFullMask, w, k); if (lane >= k) w += y; }
            s_w[lane] = w;
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        const uint64_t incl = x + (warp ? s_w[warp - 1] : 0);
        if (i < nchunks) chunks[i].offset = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) result[0] = s_carry - base;
}

// ------------------------------------------------------------------------------------------
// K3: bit packing
// ------------------------------------------------------------------------------------------
constexpr int kPackThreads = 256;
constexpr int kStageWords = (kPackThreads * 48 + 64) / 32 + 4;

struct Packer {
    uint32_t* stage;                                            // shared staging window, bit 0 = first unwritten bit of byte `bytepos`
    uint32_t* warp_sums;                                        // shared, kPackThreads/32 entries
    uint8_t* dst;                                               // chunk output
    uint64_t bytepos;                                           // bytes already written to dst
    uint32_t carry_bits;                                        // bits pending in stage[0] (0..7)

    // every thread contributes (value, nbits <= 48); bits are appended in thread order
    __device__ void round(uint64_t value, uint32_t nbits)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t x = nbits;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) { const uint32_t y = __shfl_up_sync(kFullMask, x, k); if (lane >= k) x += y; }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kPackThreads / 32; w++) { const uint32_t s = warp_sums[w]; if (w < warp) before += s; total += s; }
        const uint32_t off = carry_bits + before + x - nbits;
        if (nbits) {
            const uint32_t sh = off & 31, wi = off >> 5;
            const uint64_t lo = value << sh;
            atomicOr(&stage[wi], (uint32_t)lo);
            const uint32_t mid = (uint32_t)(lo >> 32);
            if (mid) atomicOr(&stage[wi + 1], mid);
            if (sh) { const uint32_t hi = (uint32_t)(value >> (64 - sh)); if (hi) atomicOr(&stage[wi + 2], hi); }
        }
        __syncthreads();
        const uint32_t tbits = carry_bits + total, nbytes = tbits >> 3;
        const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
        for (uint32_t i = threadIdx.x; i < nbytes; i += kPackThreads) dst[bytepos + i] = sb[i];
        const uint32_t tail = (tbits & 7) ? sb[nbytes] : 0u;
        __syncthreads();
        const uint32_t used_words = (tbits + 31) / 32 + 1;
        for (uint32_t i = threadIdx.x; i < used_words && i < (uint32_t)kStageWords; i += kPackThreads) stage[i] = i == 0 ? tail : 0u;
        __syncthreads();
        bytepos += nbytes; carry_bits = tbits & 7;
    }
};

__global__ void __launch_bounds__(kPackThreads)
k_huff_pack(const uint8_t* __restrict__ src, uint64_t n, const uint32_t* __restrict__ tok,
            const BlockMeta* __restrict__ blk, const uint32_t* __restrict__ blk_codes, const uint32_t* __restrict__ blk_hdr,
            const ChunkMeta* __restrict__ chunks, uint8_t* __restrict__ out, uint64_t cap, int last_is_final,
            int force_mark, uint32_t* __restrict__ err)
{
    __shared__ uint32_t s_stage[kStageWords];
    __shared__ uint32_t s_sums[kPackThreads / 32];
    __shared__ uint32_t s_codes[kHistSize];
    const uint64_t c = blockIdx.x;
    const uint64_t nchunks = gridDim.x;
    const ChunkMeta cm = chunks[c];
    if (cm.offset + cm.bytes > cap) return;                     // caller reports Z_BUF_ERROR from the total
    const uint64_t cbeg = c * kChunk;
    const uint32_t clen = (uint32_t)min((uint64_t)kChunk, n - cbeg);
    const bool final_chunk = last_is_final && c == nchunks - 1;
    const bool mark = force_mark && !last_is_final && c == nchunks - 1;
    uint8_t* dst = out + cm.offset;
    const uint8_t* in = src + cbeg;

    if (cm.stored) {                                            // deflate_stored / _tr_stored_block (trees.c:867)
        uint32_t done = 0; uint64_t o = 0;
        while (done < clen) {
            const uint32_t len = min(65535u, clen - done);
            const bool lastb = final_chunk && done + len == clen;
            if (threadIdx.x == 0) {
                dst[o] = lastb ? 1 : 0;
                dst[o + 1] = (uint8_t)len; dst[o + 2] = (uint8_t)(len >> 8);
                dst[o + 3] = (uint8_t)~len; dst[o + 4] = (uint8_t)(~len >> 8);
            }
            for (uint32_t i = threadIdx.x; i < len; i += kPackThreads) dst[o + 5 + i] = in[done + i];
            o += 5 + len; done += len;
        }
        if (mark && threadIdx.x == 0) { dst[o] = 0; dst[o + 1] = 0; dst[o + 2] = 0; dst[o + 3] = 0xff; dst[o + 4] = 0xff; }
        return;
    }

    for (int i = threadIdx.x; i < kStageWords; i += kPackThreads) s_stage[i] = 0;
    __syncthreads();
    Packer pk{s_stage, s_sums, dst, 0, 0};
    bool ends_stored = false;
    for (uint32_t b = 0; b < cm.nblocks; b++) {
        const BlockMeta bm = blk[c * kMaxBlocks + b];
        const uint32_t bfinal = (final_chunk && b == cm.nblocks - 1) ? 1u : 0u;
        if (bm.type == 0) {
            // 3 header bits, pad to a byte, LEN, NLEN, then the raw bytes
            const uint32_t pad = (8 - ((pk.carry_bits + 3) & 7)) & 7;
            uint64_t v = 0; uint32_t nb = 0;
            if (threadIdx.x == 0) { v = bfinal; nb = 3 + pad; }
            else if (threadIdx.x == 1) { v = (bm.in_len & 0xffffu) | ((~bm.in_len & 0xffffu) << 16); nb = 32; }
            pk.round(v, nb);
            for (uint32_t i = threadIdx.x; i < bm.in_len; i += kPackThreads) dst[pk.bytepos + i] = in[bm.in_start + i];
            pk.bytepos += bm.in_len;
            ends_stored = true;
            continue;
        }
        ends_stored = false;
        for (int i = threadIdx.x; i < kHistSize; i += kPackThreads) s_codes[i] = blk_codes[(c * kMaxBlocks + b) * kHistSize + i];
        {   // block header: 3 bits, then the serialised trees of a dynamic block
            const uint32_t hw = bm.type == 2 ? (bm.hdr_bits + 31) / 32 : 0;
            const uint32_t* hdr = blk_hdr + (c * kMaxBlocks + b) * kHdrWords;
            uint64_t v = 0; uint32_t nb = 0;
            if (threadIdx.x == 0) { v = bfinal | (bm.type << 1); nb = 3; }
            else if (threadIdx.x <= hw) {
                const uint32_t i = threadIdx.x - 1;
                v = hdr[i]; nb = (i == hw - 1 && (bm.hdr_bits & 31)) ? (bm.hdr_bits & 31) : 32;
            }
            pk.round(v, nb);                                    // also publishes s_codes (barrier inside)
        }
        const uint32_t* t = tok + cbeg + bm.tok_start;
        for (uint32_t i0 = 0; i0 <= bm.tok_count; i0 += kPackThreads) {
            const uint32_t i = i0 + threadIdx.x;
            uint64_t v = 0; uint32_t nb = 0;
            if (i < bm.tok_count) {
                const uint32_t tk = t[i], dist = tk >> 16;
                if (dist == 0) {
                    const uint32_t e = s_codes[tk & 0xff];
                    v = e & 0xffffu; nb = e >> 16;
                } else {
                    const uint32_t l = tk & 0xff, lc = len_code(l), le = len_extra_bits(lc);
                    const uint32_t e = s_codes[257 + lc];
                    v = e & 0xffffu; nb = e >> 16;
                    if (le) { v |= (uint64_t)(l & ((1u << le) - 1u)) << nb; nb += le; }
                    const uint32_t d = dist - 1, dc = dist_code(d), de = dist_extra_bits(dc);
                    const uint32_t f = s_codes[288 + dc];
                    v |= (uint64_t)(f & 0xffffu) << nb; nb += f >> 16;
                    if (de) { v |= (uint64_t)(d & ((1u << de) - 1u)) << nb; nb += de; }
                }
            } else if (i == bm.tok_count) {
                const uint32_t e = s_codes[256];                // end of block
                v = e & 0xffffu; nb = e >> 16;
            }
            pk.round(v, nb);
        }
    }
    // end of chunk: final -> pad; otherwise empty stored block unless already byte-aligned by a stored block
    if (final_chunk || (ends_stored && !mark)) {
        const uint32_t pad = (8 - (pk.carry_bits & 7)) & 7;
        pk.round(0, threadIdx.x == 0 ? pad : 0);
    } else {
        const uint32_t pad = (8 - ((pk.carry_bits + 3) & 7)) & 7;
        uint64_t v = 0; uint32_t nb = 0;
        if (threadIdx.x == 0) nb = 3 + pad;
        else if (threadIdx.x == 1) { v = 0xffff0000u; nb = 32; }
        pk.round(v, nb);
    }
    if (threadIdx.x == 0 && (pk.bytepos != cm.bytes || pk.carry_bits != 0)) atomicAdd(err, 1u);
}

// ------------------------------------------------------------------------------------------
// stream framing: header and trailer (deflate.c:577-650, 832-850)
// ------------------------------------------------------------------------------------------
__global__ void k_frame(uint8_t* __restrict__ out, uint64_t cap, uint64_t hdr_len, const uint64_t* __restrict__ payload,
                        const uint32_t* __restrict__ sums, uint64_t n, int level, int wrap, int flags, uint64_t* __restrict__ total_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint64_t pos = hdr_len + payload[0];
    uint8_t trailer[8]; int tl = 0;
    if (!(flags & (ZB200_DEFLATE_NO_TRAILER | ZB200_DEFLATE_NOT_LAST))) {
        if (wrap == ZB200_WRAP_ZLIB) {
            const uint32_t a = sums[1];
            trailer[0] = a >> 24; trailer[1] = a >> 16; trailer[2] = a >> 8; trailer[3] = a; tl = 4;
        } else if (wrap == ZB200_WRAP_GZIP) {
            const uint32_t cr = sums[0];
            for (int i = 0; i < 4; i++) trailer[i] = cr >> (8 * i);
            for (int i = 0; i < 4; i++) trailer[4 + i] = (uint8_t)(n >> (8 * i));
            tl = 8;
        }
    }
    total_out[0] = pos + tl;
    if (pos + tl > cap) return;
    if (hdr_len) {
        if (wrap == ZB200_WRAP_ZLIB) {
            const uint32_t fl = level < 2 ? 0 : level < 6 ? 1 : level == 6 ? 2 : 3;
            uint32_t h = (0x78u << 8) | (fl << 6);
            h += 31 - h % 31;
            out[0] = h >> 8; out[1] = h;
        } else if (wrap == ZB200_WRAP_GZIP) {
            const uint8_t g[10] = {31, 139, 8, 0, 0, 0, 0, 0, (uint8_t)(level == 9 ? 2 : level < 2 ? 4 : 0), 3};
            for (int i = 0; i < 10; i++) out[i] = g[i];
        }
    }
    for (int i = 0; i < tl; i++) out[pos + i] = trailer[i];
}

// empty input: a lone final fixed block with just the end-of-block code = 03 00 (what the reference emits)
__global__ void k_empty_payload(uint8_t* out, uint64_t cap, uint64_t at, int not_last, int force_mark, uint64_t* payload)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (not_last) {
        payload[0] = force_mark ? 5 : 0;
        if (force_mark && at + 5 <= cap) { out[at] = 0; out[at + 1] = 0; out[at + 2] = 0; out[at + 3] = 0xff; out[at + 4] = 0xff; }
        return;
    }
    payload[0] = 2;
    if (at + 2 <= cap) { out[at] = 3; out[at + 1] = 0; }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// d_buf = [dict bytes][n source bytes], contiguous in device memory.  d_res: u64[0] total bytes,
// u32 at +8.. : crc, adler, error count.
int deflate_launch(Ctx* c, const uint8_t* d_buf, uint64_t dict, uint64_t n, uint8_t* d_out, uint64_t cap, int level,
                   int wrap, int flags, uint64_t* d_total, uint32_t* d_sums, uint32_t* d_err, cudaStream_t s)
{
    static bool attr = false;
    if (!attr) {
        ZB_CUDA(cudaFuncSetAttribute(k_lz_link, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << kHashBits) * 2));
        attr = true;
    }
    if (level < 0) level = 6;
    LevelCfg cfg = h_levels[level];
    const int strategy = (flags >> 8) & 7;                      // Z_FILTERED 1, Z_HUFFMAN_ONLY 2, Z_RLE 3, Z_FIXED 4
    const int force_mark = (flags & ZB200I_DEFLATE_FORCE_MARK) ? 1 : 0;
    if (strategy == 2 && cfg.kind != 0) { cfg.chain = 0; cfg.kind = 1; }   // literals only (deflate.c:1490)
    const uint64_t total = dict + n;
    const uint64_t nchunks = (n + kChunk - 1) / kChunk;
    const int last_is_final = (flags & ZB200_DEFLATE_NOT_LAST) ? 0 : 1;
    const uint64_t hdr_len = (flags & ZB200_DEFLATE_NO_HEADER) ? 0 : wrap == ZB200_WRAP_ZLIB ? 2 : wrap == ZB200_WRAP_GZIP ? 10 : 0;
    const uint8_t* d_src = d_buf + dict;
    int rc;

    // checksums of the uncompressed data (read_buf, deflate.c:956-981)
    if ((rc = checksum_launch(c, d_src, n, d_sums, s)) != 0) return rc;
    ZB_CUDA(cudaMemsetAsync(d_err, 0, 4, s));
    if ((rc = c->ws[1].ensure(16)) != 0) return rc;
    uint64_t* d_payload = c->ws[1].as<uint64_t>();

    if (nchunks == 0) {
        ZB_LAUNCH(k_empty_payload, 1, 32, 0, s, d_out, cap, hdr_len, !last_is_final, force_mark, d_payload);
    } else {
        if ((rc = c->ws[2].ensure(nchunks * sizeof(ChunkMeta))) != 0) return rc;
        ChunkMeta* d_chunks = c->ws[2].as<ChunkMeta>();
        const uint64_t nslots = nchunks * kMaxBlocks;
        BlockMeta* d_blk = nullptr; uint32_t *d_hist = nullptr, *d_codes = nullptr, *d_hdr = nullptr, *d_tok = nullptr;
        if (cfg.kind != 0) {
            if ((rc = c->ws[3].ensure(total * 2 + 64)) != 0) return rc;          // dist16
            if ((rc = c->ws[4].ensure(n * 4 + 64)) != 0) return rc;              // per-position matches
            if ((rc = c->ws[5].ensure(n * 4 + 64)) != 0) return rc;              // tokens
            if ((rc = c->ws[6].ensure(nslots * sizeof(BlockMeta))) != 0) return rc;
            if ((rc = c->ws[7].ensure(nslots * kHistSize * 4)) != 0) return rc;
            if ((rc = c->ws[8].ensure(nslots * kHistSize * 4)) != 0) return rc;
            if ((rc = c->ws[9].ensure(nslots * kHdrWords * 4)) != 0) return rc;
            uint16_t* d_dist = c->ws[3].as<uint16_t>();
            uint32_t* d_mt = c->ws[4].as<uint32_t>();
            d_tok = c->ws[5].as<uint32_t>();
            d_blk = c->ws[6].as<BlockMeta>();
            d_hist = c->ws[7].as<uint32_t>(); d_codes = c->ws[8].as<uint32_t>(); d_hdr = c->ws[9].as<uint32_t>();
            const unsigned nseg = (unsigned)((total + kChunk - 1) / kChunk);
            if (cfg.chain == 0) {
                ZB_CUDA(cudaMemsetAsync(d_mt, 0, n * 4, s));
            } else {
                ZB_LAUNCH(k_lz_link, nseg, 32, (1 << kHashBits) * 2, s, d_buf, total, d_dist);
                ZB_LAUNCH(k_lz_match, (unsigned)((n + 255) / 256), 256, 0, s, d_buf, total, dict, d_dist, d_mt, (int)cfg.chain, (int)cfg.nice);
            }
            ZB_LAUNCH(k_lz_parse, (unsigned)((nchunks + kParseWarps - 1) / kParseWarps), kParseWarps * 32, 0, s, d_src, n, d_mt,
                      d_tok, d_blk, d_hist, d_chunks, cfg.kind, (uint32_t)cfg.lazy);
            ZB_LAUNCH(k_huff_build, (unsigned)((nslots + kCodeWarps - 1) / kCodeWarps), kCodeWarps * 32, 0, s, nchunks, d_chunks,
                      d_blk, d_hist, d_codes, d_hdr, strategy == 4 ? 1 : 0);
        } else {
            ZB_CUDA(cudaMemsetAsync(d_chunks, 0, nchunks * sizeof(ChunkMeta), s));
        }
        ZB_LAUNCH(k_plan, (unsigned)((nchunks + 255) / 256), 256, 0, s, nchunks, n, d_chunks, d_blk, last_is_final, cfg.kind == 0 ? 1 : 0, force_mark);
        ZB_LAUNCH(k_scan, 1, 1024, 0, s, nchunks, d_chunks, hdr_len, d_payload);
        ZB_LAUNCH(k_huff_pack, (unsigned)nchunks, kPackThreads, 0, s, d_src, n, d_tok, d_blk, d_codes, d_hdr, d_chunks, d_out, cap,
                  last_is_final, force_mark, d_err);
    }
    ZB_LAUNCH(k_frame, 1, 32, 0, s, d_out, cap, hdr_len, d_payload, d_sums, n, level, wrap, flags, d_total);
    ZB_CHECK_LAUNCH();
    return 0;
}

}  // namespace zb

using namespace zb;

ZB_API int zb200_deflate_shard(const void* src, size_t src_len, const void* dict, size_t dict_len, void* dst,
                               size_t* dst_len, int level, int wrap, int flags, uint32_t* crc, uint32_t* adler, void* stream)
{
    int rc = ensure_init();
    if (rc) return rc;
    if (!dst_len || level < -1 || level > 9 || wrap < 0 || wrap > 2 || (src_len && !src) || dict_len > kWindow) {
        set_error("zb200_deflate: bad argument");
        return ZB_STREAM_ERROR;
    }
    Ctx* c = ctx_acquire((cudaStream_t)stream);
    if (!c) return ZB_MEM_ERROR;
    cudaStream_t s = pick_stream(c, stream);
    const size_t cap = *dst_len;
    do {
        // bring [dict][src] to one contiguous device range
        const uint8_t* d_buf;
        if (dict_len == 0) {
            d_buf = to_device(c, src, src_len, s, &rc);
            if (rc) break;
        } else if (classify(src) == kDevice && classify(dict) == kDevice &&
                   (const uint8_t*)dict + dict_len == (const uint8_t*)src) {
            d_buf = (const uint8_t*)dict;
        } else {
            if ((rc = c->in.ensure(dict_len + src_len + 64)) != 0) break;
            cudaError_t e = cudaMemcpyAsync(c->in.p, dict, dict_len, cudaMemcpyDefault, s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c->in.as<uint8_t>() + dict_len, src, src_len, cudaMemcpyDefault, s);
            if (e != cudaSuccess) { set_error("input staging failed: %s", cudaGetErrorString(e)); rc = ZB_STREAM_ERROR; break; }
            d_buf = c->in.as<uint8_t>();
        }
        const bool dst_on_host = classify(dst) != kDevice;
        uint8_t* d_out = (uint8_t*)dst;
        if (dst_on_host) {
            if ((rc = c->out.ensure(cap + 16)) != 0) break;
            d_out = c->out.as<uint8_t>();
        }
        if ((rc = c->small.ensure(256)) != 0) break;
        if ((rc = c->ensure_pinned(256)) != 0) break;
        uint64_t* d_total = c->small.as<uint64_t>();
        uint32_t* d_sums = c->small.as<uint32_t>() + 2;
        uint32_t* d_err = c->small.as<uint32_t>() + 4;
        if ((rc = deflate_launch(c, d_buf, dict_len, src_len, d_out, cap, level, wrap, flags, d_total, d_sums, d_err, s)) != 0) break;
        cudaError_t e = cudaMemcpyAsync(c->pinned, c->small.p, 32, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { set_error("deflate failed: %s", cudaGetErrorString(e)); rc = ZB_STREAM_ERROR; break; }
        const uint64_t total = *(const uint64_t*)c->pinned;
        const uint32_t* r32 = (const uint32_t*)c->pinned;
        if (crc) *crc = r32[2];
        if (adler) *adler = r32[3];
        if (r32[4] != 0) { set_error("internal error: %u chunks packed to a size other than planned", r32[4]); rc = ZB_STREAM_ERROR; break; }
        *dst_len = (size_t)total;
        if (total > cap) { rc = ZB_BUF_ERROR; set_error("output buffer too small: need %llu, have %zu", (unsigned long long)total, cap); break; }
        if (dst_on_host && total) {
            e = cudaMemcpyAsync(dst, d_out, total, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = ZB_STREAM_ERROR; break; }
        }
    } while (0);
    ctx_release(c, s);
    return rc;
}

ZB_API int zb200_deflate(const void* src, size_t src_len, void* dst, size_t* dst_len, int level, int wrap, void* stream)
{
    return zb200_deflate_shard(src, src_len, nullptr, 0, dst, dst_len, level, wrap, 0, nullptr, nullptr, stream);
}
