// inflate_core.cuh — RFC 1950/1951 decoder and the CompaCT op-stream decoder as host/device functions
// (the kernels in cct_decode.cu and the CPU harness in tests/hostsim run the same code).
//
// Restates: `zlib.decompress(file_bytes[13:])` (reference src/codec/core.py:420-421; any conformant
// inflate reproduces it) and the decode loop (core.py:453-520) in the parallel-friendly
// formulation of SURVEY Appendix A.5: ops -> deltas -> running sum; JUMPs -> unit table -> scatter.
#pragma once
#include "cct_common.cuh"

namespace cct {

constexpr int kLitBits = 10;  // primary lookup width, literal/length codes
constexpr int kInfSlack = 64; // spare symbol slots the symbol buffer must have beyond the largest valid count
constexpr int kDstBits = 9;   // primary lookup width, distance codes

struct InflateTables {  // per-stream decode tables (shared memory on the device)
    uint16_t lit[1 << kLitBits];  // len << 12 | symbol ; 0 = code longer than kLitBits (slow path)
    uint16_t dst[1 << kDstBits];
    uint16_t lcount[16], dcount[16];   // canonical: number of codes of each length
    uint16_t lsym[288], dsym[32];      // symbols ordered by (length, symbol)
    uint32_t pad;                      // size = 945 words: consecutive tables start in different smem banks
};

// RFC 1951 §3.2.5 length / distance bases as closed forms
CCT_HD int len_extra_bits_inf(int idx) { return (idx < 8 || idx == 28) ? 0 : (idx >> 2) - 1; }
CCT_HD int len_base_inf(int idx) {
    if (idx < 8) return 3 + idx;
    if (idx == 28) return 258;
    return 3 + ((4 + (idx & 3)) << ((idx >> 2) - 1));
}
CCT_HD int dist_base_inf(int ds) { return ds < 4 ? 1 + ds : 1 + ((2 + (ds & 1)) << ((ds >> 1) - 1)); }

struct BitReader {
    // LSB-first bit buffer over in[0..len).  After align_input() the read position is 4-byte aligned, so a
    // refill is one 32-bit load; only the last few bytes of the stream go through the byte path.  Reading past
    // the end yields zero bits and drives nbits negative, which the callers check at block boundaries.
    const uint8_t *in; int len, pos; uint64_t buf; int nbits; bool bad;
    CCT_HD void align_input() {
        while (pos < len && ((uintptr_t)(in + pos) & 3u)) { buf |= (uint64_t)in[pos++] << nbits; nbits += 8; }
    }
    CCT_HD void refill() {
        if (nbits > 32) return;
        if (pos + 4 <= len) {
#if defined(CCT_STATS) && !defined(__CUDA_ARCH__)
            if ((uintptr_t)(in + pos) & 3u) abort();  // CPU harness: the GPU would fault on this load
#endif
            buf |= (uint64_t)(*reinterpret_cast<const uint32_t *>(in + pos)) << nbits;
            pos += 4; nbits += 32;
            return;
        }
        while (nbits <= 56 && pos < len) { buf |= (uint64_t)in[pos++] << nbits; nbits += 8; }
    }
    CCT_HD void seek(uint32_t bitpos) {  // stand at an arbitrary bit of the stream
        pos = (int)(bitpos >> 3); buf = 0; nbits = 0; bad = false;
        align_input(); refill(); drop((int)(bitpos & 7u));
    }
    CCT_HD uint32_t tell() const { return 8u * (uint32_t)pos - (uint32_t)nbits; }
    CCT_HD uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    CCT_HD void drop(int n) { buf >>= n; nbits -= n; if (nbits < 0) bad = true; }
    CCT_HD uint32_t take(int n) { uint32_t v = peek(n); drop(n); return v; }
};

// Build primary table + canonical arrays from code lengths.  Returns false for an over-subscribed set.
CCT_HDN inline bool build_decode_table(const uint8_t *lens, int nsym, int prim_bits, uint16_t *prim, uint16_t *count,
                                       uint16_t *symbols) {
    uint16_t offs[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < nsym; i++) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int i = 0; i < nsym; i++) if (lens[i]) symbols[offs[lens[i]]++] = (uint16_t)i;
    for (int i = 0; i < (1 << prim_bits); i++) prim[i] = 0;
    // canonical codes, shortest first
    uint32_t code = 0; int idx = 0;
    for (int l = 1; l <= prim_bits; l++) {
        for (int c = 0; c < count[l]; c++, idx++, code++) {
            uint32_t rev = 0;
            for (int b = 0; b < l; b++) rev |= ((code >> b) & 1u) << (l - 1 - b);
            uint16_t e = (uint16_t)((l << 12) | symbols[idx]);
            for (uint32_t k = rev; k < (1u << prim_bits); k += (1u << l)) prim[k] = e;
        }
        code <<= 1;
    }
    return true;
}

// Decode one symbol; drops its bits.  Returns -1 on an invalid code.
CCT_HD int decode_symbol(BitReader &br, const uint16_t *prim, int prim_bits, const uint16_t *count, const uint16_t *symbols) {
    uint16_t e = prim[br.peek(prim_bits)];
    if (e) { br.drop(e >> 12); return e & 0xFFF; }
    // slow path: canonical bit-by-bit (codes longer than prim_bits)
    int code = 0, first = 0, index = 0;
    uint64_t b = br.buf;
    for (int l = 1; l <= 15; l++) {
        code |= (int)(b & 1); b >>= 1;
        int c = count[l];
        if (code - c < first) { br.drop(l); return symbols[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// Pass 1 of inflate: zlib stream -> LZ77 symbols (same packing as the encoder: bits 0..7 literal byte or len-3,
// bits 8..23 distance, 0 = literal).  Nothing but the bit reader and the code tables is touched, so one
// thread per stream can run it with its tables in shared memory.  `inflate_step` handles everything outside
// the symbol loop (block header + tables, stored bytes, trailer) as one step of a small state machine;
// `inflate_symbols` runs the machine with the symbol phase as a tight loop.
enum { kInfHeader = 0, kInfSymbol = 1, kInfStored = 2, kInfTrailer = 3, kInfDone = 4, kInfError = 5 };
struct InflateState {
    BitReader br;
    int phase, ns, op, stored_left;
    bool last;
    uint32_t adler_want;
};

CCT_HD void inflate_begin(InflateState &st, const uint8_t *in, int in_len) {
    st.br = BitReader{in, in_len - 4, 2, 0, 0, false};
    st.br.align_input();
    st.ns = 0; st.op = 0; st.stored_left = 0; st.last = false; st.adler_want = 1;
    st.phase = kInfHeader;
    if (in_len < 6 || (in[0] & 0x0F) != 8 || (((uint32_t)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) st.phase = kInfError;
}

// Block header + code tables (RFC 1951 3.2.3 - 3.2.7).  Returns the next phase.
CCT_HDN inline int inflate_header(InflateState &st, InflateTables &tb) {
    BitReader &br = st.br;
    uint8_t lens[320];
    br.refill();
    st.last = br.take(1);
    const int type = (int)br.take(2);
    if (type == 0) {
        br.drop(br.nbits & 7);
        br.refill();
        const uint32_t len = br.take(16), nlen = br.take(16);
        if (br.bad || (len ^ 0xFFFFu) != nlen) return kInfError;
        st.stored_left = (int)len;
        return kInfStored;
    }
    if (type == 3) return kInfError;
    if (type == 1) {
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        build_decode_table(lens, 288, kLitBits, tb.lit, tb.lcount, tb.lsym);
        for (int i = 0; i < 30; i++) lens[i] = 5;
        build_decode_table(lens, 30, kDstBits, tb.dst, tb.dcount, tb.dsym);
        return kInfSymbol;
    }
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
    if (nlen > 286 || ndist > 30) return kInfError;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (int i = 0; i < ncode; i++) { br.refill(); cl[order[i]] = (uint8_t)br.take(3); }
    if (!build_decode_table(cl, 19, 7, tb.dst, tb.dcount, tb.dsym)) return kInfError;  // borrows the distance table
    int idx = 0;
    while (idx < nlen + ndist) {
        br.refill();
        const int sym = decode_symbol(br, tb.dst, 7, tb.dcount, tb.dsym);
        if (sym < 0 || br.bad) return kInfError;
        if (sym < 16) lens[idx++] = (uint8_t)sym;
        else {
            int rep, val = 0;
            if (sym == 16) { if (idx == 0) return kInfError; val = lens[idx - 1]; rep = 3 + (int)br.take(2); }
            else if (sym == 17) rep = 3 + (int)br.take(3);
            else rep = 11 + (int)br.take(7);
            if (idx + rep > nlen + ndist) return kInfError;
            while (rep--) lens[idx++] = (uint8_t)val;
        }
    }
    if (lens[256] == 0) return kInfError;
    uint8_t dl[32];
    for (int i = 0; i < ndist; i++) dl[i] = lens[nlen + i];
    if (!build_decode_table(lens, nlen, kLitBits, tb.lit, tb.lcount, tb.lsym)) return kInfError;
    if (!build_decode_table(dl, ndist, kDstBits, tb.dst, tb.dcount, tb.dsym)) return kInfError;
    return kInfSymbol;
}

// One literal/length(+distance) symbol of the current block; the hot path of pass 1.  Leaves kInfSymbol at the
// end of the block or on an error.  Bits past the end of the input read as zero and are caught at the block end.
CCT_HD void inflate_one_symbol(InflateState &st, const InflateTables &tb, uint32_t *syms, int sym_cap, int out_cap) {
    BitReader &br = st.br;
    if (st.ns >= sym_cap) { st.phase = kInfError; return; }
    br.refill();
    const uint32_t e = tb.lit[br.buf & ((1u << kLitBits) - 1)];
    int sym;
    if (e) {
        const int l = (int)(e >> 12);
        br.buf >>= l; br.nbits -= l;
        sym = (int)(e & 0xFFF);
    } else {
        sym = decode_symbol(br, tb.lit, kLitBits, tb.lcount, tb.lsym);
        if (sym < 0) { st.phase = kInfError; return; }
    }
    if (sym < 256) { syms[st.ns++] = (uint32_t)sym; st.op++; return; }
    if (sym == 256) {
        st.phase = (br.nbits < 0 || br.bad || st.op > out_cap) ? kInfError : (st.last ? kInfTrailer : kInfHeader);
        return;
    }
    sym -= 257;
    if (sym >= 29) { st.phase = kInfError; return; }
    const int len = len_base_inf(sym) + (int)br.take(len_extra_bits_inf(sym));
    const int ds = decode_symbol(br, tb.dst, kDstBits, tb.dcount, tb.dsym);
    if (ds < 0 || ds >= 30) { st.phase = kInfError; return; }
    br.refill();
    const int dist = dist_base_inf(ds) + (int)br.take(ds < 4 ? 0 : (ds >> 1) - 1);
    if (dist > st.op) { st.phase = kInfError; return; }
    syms[st.ns++] = (uint32_t)(len - 3) | ((uint32_t)dist << 8);
    st.op += len;
}

CCT_HD void inflate_step(InflateState &st, InflateTables &tb, uint32_t *syms, int sym_cap, int out_cap) {
    BitReader &br = st.br;
    if (st.phase == kInfSymbol) {
        inflate_one_symbol(st, tb, syms, sym_cap, out_cap);
    } else if (st.phase == kInfHeader) {
        st.phase = inflate_header(st, tb);
    } else if (st.phase == kInfStored) {  // stored bytes become literal symbols, a few per step
        int k = st.stored_left < 8 ? st.stored_left : 8;
        if (st.ns + k > sym_cap || st.op + k > out_cap) { st.phase = kInfError; return; }
        st.stored_left -= k; st.op += k;
        for (; k > 0; k--) {
            uint8_t c;
            if (br.nbits >= 8) c = (uint8_t)br.take(8);
            else { if (br.pos >= br.len) { st.phase = kInfError; return; } c = br.in[br.pos++]; }
            syms[st.ns++] = c;
        }
        if (st.stored_left == 0) { br.align_input(); st.phase = st.last ? kInfTrailer : kInfHeader; }  // bytes were read one by one
    } else if (st.phase == kInfTrailer) {  // the Adler-32 trailer follows the last block at the next byte boundary
        if (br.bad) { st.phase = kInfError; return; }
        br.drop(br.nbits & 7);
        if (br.pos - (br.nbits >> 3) != br.len) { st.phase = kInfError; return; }
        const uint8_t *t = br.in + br.len;
        st.adler_want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
        st.phase = kInfDone;
    }
}

// Whole stream in one call: the same machine with the symbol phase as a tight inner loop (literal fast path,
// one bounds check per symbol).  This is what a thread runs when it does not share its warp with other streams.
// Returns the number of symbols or -1.
CCT_HDN inline int inflate_symbols(const uint8_t *in, int in_len, uint32_t *syms, int sym_cap, int out_cap,
                                   InflateTables &tb, int *out_len, uint32_t *adler_want) {
    InflateState st;
    inflate_begin(st, in, in_len);
    BitReader &br = st.br;
    while (st.phase < kInfDone) {
        if (st.phase != kInfSymbol) { inflate_step(st, tb, syms, sym_cap, out_cap); continue; }
        // hot loop: bit buffer, output cursor and the running output length live in locals; the output length
        // is (symbols written) + (sum of len-1 over matches), so a literal costs no bookkeeping beyond the store
        uint64_t buf = br.buf;
        int nbits = br.nbits, pos = br.pos;
        const uint8_t *in_ = br.in;
        const int in_len_ = br.len;
        uint32_t *p = syms + st.ns;
        uint32_t *const p0 = p, *const pend = syms + sym_cap;
        int extra = 0;  // sum of (len - 1) over the matches of this block
        const int op0 = st.op;
        int next = kInfError;
        for (;;) {
            if (nbits <= 32) {
                // room for the symbols of the next 32+ bits is checked here, not per symbol: at most 32 symbols
                // are decoded between two refills (kInfSlack = 64 spare slots)
                if (pend - p < kInfSlack) break;
                if (pos + 4 <= in_len_) { buf |= (uint64_t)(*reinterpret_cast<const uint32_t *>(in_ + pos)) << nbits; pos += 4; nbits += 32; }
                else while (nbits <= 56 && pos < in_len_) { buf |= (uint64_t)in_[pos++] << nbits; nbits += 8; }
            }
            const uint32_t e = tb.lit[(uint32_t)buf & ((1u << kLitBits) - 1)];
            int sym;
            if (e) {
                const int l = (int)(e >> 12);
                buf >>= l; nbits -= l;
                sym = (int)(e & 0xFFF);
            } else {
                br.buf = buf; br.nbits = nbits; br.pos = pos;
                sym = decode_symbol(br, tb.lit, kLitBits, tb.lcount, tb.lsym);
                buf = br.buf; nbits = br.nbits;
                if (sym < 0) break;
            }
            if (sym < 256) { *p++ = (uint32_t)sym; continue; }  // literal
            if (sym == 256) {
                const int op = op0 + (int)(p - p0) + extra;
                next = (nbits < 0 || op > out_cap) ? kInfError : (st.last ? kInfTrailer : kInfHeader);
                break;
            }
            sym -= 257;
            if (sym >= 29 || pend - p < kInfSlack) break;  // (the refills below bypass the capacity check above)
            br.buf = buf; br.nbits = nbits; br.pos = pos;
            const int len = len_base_inf(sym) + (int)br.take(len_extra_bits_inf(sym));
            const int ds = decode_symbol(br, tb.dst, kDstBits, tb.dcount, tb.dsym);
            if (ds < 0 || ds >= 30) break;
            br.refill();
            const int dist = dist_base_inf(ds) + (int)br.take(ds < 4 ? 0 : (ds >> 1) - 1);
            buf = br.buf; nbits = br.nbits; pos = br.pos;
            if (dist > op0 + (int)(p - p0) + extra) break;
            *p++ = (uint32_t)(len - 3) | ((uint32_t)dist << 8);
            extra += len - 1;
        }
        br.buf = buf; br.nbits = nbits; br.pos = pos;
        st.op = op0 + (int)(p - p0) + extra;
        st.ns = (int)(p - syms);
        st.phase = next;
    }
    if (st.phase != kInfDone || br.bad) return -1;
    *out_len = st.op; *adler_want = st.adler_want;
    return st.ns;
}

// ------------------------------------------------------------------------------------------------
// Lane-parallel Huffman decode (the form k_inflate_sym runs; tests/hostsim steps the same scalar pieces)
// ------------------------------------------------------------------------------------------------
// One stream is decoded by the kSubLanes threads of a CTA.  Blocks are sequential (a block's header is found only
// when the previous block ends), but inside a block the bit string is cut into rounds of kSubLanes sub-sequences
// of kSubBits bits, one per lane.  Lane 0 starts on a true symbol boundary; the others start speculatively at
// their sub-sequence boundary.  Huffman codes self-synchronise: a decoder started at a wrong bit falls into step with
// the true symbol sequence after a few symbols.  So every lane decodes to the end of its sub-sequence and reports
// where its last symbol ended; lane i+1 then restarts from lane i's end.  A lane is CONFIRMED once it started
// where its confirmed predecessor ended; rounds repeat (normally twice) until every lane up to the end of the
// block is confirmed, then each lane decodes once more writing its symbols at the offset given by a scan of the
// symbol counts.  The reader is stateless (two staged words + a funnel shift per peek), so lanes stay in lock step.
#ifndef CCT_FLIT_BITS
#define CCT_FLIT_BITS 11
#endif
constexpr int kFLitBits = CCT_FLIT_BITS, kFDstBits = 9;   // primary lookup widths (11 literal bits: codes of 11 bits are common in a CT block, and the slow path runs with two or three lanes of a warp)
constexpr int kFDstAt = 1 << kFLitBits;         // where the distance entries start in FastTables::tab
constexpr int kSubBits = 512;                  // bits per lane per round
constexpr int kSubLanes = 128;                 // sub-sequences per round = threads per stream
constexpr int kStageWords = kSubLanes * kSubBits / 32 + 8;      // one round + the longest symbol (48 bits) + peek
constexpr int kStageAlloc = kStageWords + kStageWords / 16 + 1; // one pad word per 16: lanes start 16 words apart
CCT_HD int stage_slot(int w) { return w + (w >> 4); }

// table entry: bits 0..3 code length (0 = not in the primary table), 4..5 kind, 6..9 number of extra bits, 16..31 base
// value (the literal byte, the smallest length or distance of the code)
constexpr uint32_t kEntLit = 0u << 4, kEntLen = 1u << 4, kEntEob = 2u << 4, kEntBad = 3u << 4, kEntKind = 3u << 4;

struct FastCanon {       // canonical description of one prefix code
    uint32_t lim[16];    // 31-bit left-justified exclusive upper bound of the codes of length <= l
    uint16_t first[16];  // first code of length l
    uint16_t offs[16];   // index in the sorted symbol array of the first symbol of length l
    uint16_t count[16];  // codes of length l
};
struct FastTables {      // per-stream decode state in shared memory
    uint32_t tab[(1 << kFLitBits) + (1 << kFDstBits)];  // literal/length entries, then (at kFDstAt) distance entries
    FastCanon lc, dc;
    uint32_t lsym[288], dsym[32];   // symbols ordered by (length, symbol), as table entries less the code length
    uint8_t lens[320];              // code lengths of the current block: literal/length, then distance
    uint32_t stage[kStageAlloc];    // the round's bits (padded layout); scratch while a header is parsed
};

CCT_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}

CCT_HD uint32_t lit_entry(int sym, int l) {
    if (sym < 256) return (uint32_t)l | kEntLit | ((uint32_t)sym << 16);
    if (sym == 256) return (uint32_t)l | kEntEob;
    const int idx = sym - 257;
    if (idx >= 29) return (uint32_t)l | kEntBad;
    return (uint32_t)l | kEntLen | ((uint32_t)len_extra_bits_inf(idx) << 6) | ((uint32_t)len_base_inf(idx) << 16);
}
CCT_HD uint32_t dst_entry(int sym, int l) {
    if (sym >= 30) return (uint32_t)l | kEntBad;
    return (uint32_t)l | kEntLen | ((uint32_t)(sym < 4 ? 0 : (sym >> 1) - 1) << 6) | ((uint32_t)dist_base_inf(sym) << 16);
}

// The code of `from`..`to` bits at the head of `bits` (LSB-first, as read from the stream), as a table entry;
// 0 if none.  Canonical codes, left-justified, fill consecutive intervals in order of length, so the length is
// the first l whose upper bound exceeds the (bit-reversed) head.  Requires: no code shorter than `from` matches.
// `syms` holds the symbols in canonical order, each as its table entry less the code length.
CCT_HD uint32_t canon_lookup(const FastCanon &c, const uint32_t *syms, uint32_t bits, int from, int to) {
    const uint32_t v = brev32(bits) >> 1;
    for (int l = from; l <= to; l++) {
        if (v < c.lim[l]) return syms[c.offs[l] + (int)(v >> (31 - l)) - (int)c.first[l]] | (uint32_t)l;
    }
    return 0u;
}

// Table construction, split into steps so that a warp can run the parallel ones with all lanes
// (lane, nl) = (lane id, 32) on the device and (0, 1) on the host; the caller synchronises between steps.
CCT_HD void fast_count(const uint8_t *lens, int nsym, FastCanon &c, int lane, int nl) {  // c.count zeroed by the caller
    for (int i = lane; i < nsym; i += nl) {
        const int l = lens[i];
        if (l) {
#if defined(__CUDA_ARCH__)
            atomicAdd(reinterpret_cast<unsigned int *>(&c.count[l & ~1]), (l & 1) ? 0x10000u : 1u);  // u16 pair
#else
            c.count[l]++;
#endif
        }
    }
}
// single lane: code ranges; returns false for an over-subscribed set of lengths
CCT_HDN inline bool fast_ranges(FastCanon &c) {
    int left = 1;
    uint32_t code = 0, idx = 0;
    c.count[0] = 0; c.lim[0] = 0; c.first[0] = 0; c.offs[0] = 0;
    for (int l = 1; l <= 15; l++) {
        left = (left << 1) - c.count[l];
        if (left < 0) return false;
        c.first[l] = (uint16_t)code; c.offs[l] = (uint16_t)idx;
        c.lim[l] = (code + c.count[l]) << (31 - l);
        idx += c.count[l];
        code = (code + c.count[l]) << 1;
    }
    return true;
}
// single lane: symbols ordered by (length, symbol).  `next` = 16 scratch counters.
CCT_HDN inline void fast_sort(const uint8_t *lens, int nsym, const FastCanon &c, uint32_t *syms, uint16_t *next, bool lit) {
    for (int l = 0; l < 16; l++) next[l] = c.offs[l];
    for (int i = 0; i < nsym; i++) if (lens[i]) syms[next[lens[i]]++] = lit ? lit_entry(i, 0) : dst_entry(i, 0);
}
CCT_HD void fast_fill(const FastCanon &c, const uint32_t *syms, uint32_t *prim, int prim_bits, int lane, int nl) {
    for (int t = lane; t < (1 << prim_bits); t += nl) prim[t] = canon_lookup(c, syms, (uint32_t)t, 1, prim_bits);
}

CCT_HD uint32_t stage_peek(const uint32_t *stage, uint32_t rel) {  // 32 bits starting `rel` bits into the staged words
    const int w = (int)(rel >> 5);
    const uint32_t lo = stage[stage_slot(w)], hi = stage[stage_slot(w + 1)];
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, rel & 31u);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (rel & 31u));
#endif
}

struct Span { uint32_t end; int n, extra, flag; };  // where the last symbol ended; symbols; sum of (len - 1) over matches
enum { kSpanRun = 0, kSpanEob = 1, kSpanBad = 2, kSpanFar = 3 };  // ran to the limit / end of block / invalid code or out of data / distance before the start

// Sync marks.  A lane that decodes its sub-sequence again from another start falls into step with its earlier
// list after a few symbols, and from there on the result is known.  The earlier list is remembered as one mark per
// 32-bit word of the lane's window (kSubBits + the longest symbol): the offset of the first symbol boundary in
// the word (kMarkNone: none), how many symbols start in the word and their extra bytes.  Two lists that have the same
// first boundary in a word are the same list from that boundary on.
constexpr int kMarkWords = kSubBits / 32 + 2;
constexpr uint32_t kMarkNone = 63u;
CCT_HD uint32_t mark_pack(uint32_t off, int cnt, int ext) { return off | ((uint32_t)cnt << 6) | ((uint32_t)ext << 12); }

// Decode the symbols that START in [p, limit).  Positions are in bits from the stream's aligned origin;
// `origin` is the position of staged word 0, `data_end` the end of the deflate data.  WRITE: symbols -> out[],
// `op` = bytes produced before this span (distance check).
//
// One pass of the loop reads ONE code -- literal/length, or the distance that follows a length -- and both kinds go
// through the same instructions: a table entry carries the code length, the number of extra bits and the base value,
// so "value = base + extra bits" is all there is to either.  The lanes of a warp therefore stay together whatever mix
// of literals and matches each one meets (a loop that handles a whole match in one pass runs with a dozen of its 32
// lanes active).
//
// MARKS (for the lane whose window starts at `base`): marks[w * MS] is kept up to date.  With `again`, marks[] and r
// describe the lane's previous list: the decode stops at the first word both lists enter at the same bit, and the rest
// of r comes from the marks.  Without it every mark is written.
template <bool WRITE, bool MARKS, int MS>
CCT_HD void decode_core(const uint32_t *stage, uint32_t origin, uint32_t base, uint32_t p, uint32_t limit, uint32_t data_end,
                        const FastTables &tb, uint32_t *marks, bool again, uint32_t *out, int op, Span &r) {
    int n = 0, extra = 0, flag = kSpanRun;
    int wcur = 0, n0 = 0, e0 = 0;  // the word of the current symbol; n, extra when it was entered
    uint32_t off = 0;
    if (MARKS && p < base) p = base;  // the lane before ended short of its sub-sequence: the block ends there, or it decodes again
    if (p < limit) {
        if (MARKS) {
            wcur = (int)((p - base) >> 5);
            for (int w = 0; w < wcur; w++) marks[w * MS] = kMarkNone;  // the list starts later than the previous one did
            off = (p - base) & 31u;
        }
        uint32_t mode = 0;  // 1: a length has been read, its distance is next
        int len = 0;
        bool fresh = MARKS;  // p is the first boundary of this list in word wcur
        for (;;) {
            if (MARKS && fresh) {
                fresh = false;
                if (again && (marks[wcur * MS] & 63u) == off) {
                    for (int w = wcur; w < kMarkWords; w++) {
                        const uint32_t m = marks[w * MS];
                        n += (int)((m >> 6) & 63u); extra += (int)(m >> 12);
                    }
                    r.n = n; r.extra = extra;  // r.end and r.flag stand
                    return;
                }
            }
            const uint32_t bits = stage_peek(stage, p - origin);
            uint32_t e = tb.tab[mode ? kFDstAt + (bits & ((1u << kFDstBits) - 1)) : (bits & ((1u << kFLitBits) - 1))];
            if (!e) {
                e = canon_lookup(mode ? tb.dc : tb.lc, mode ? tb.dsym : tb.lsym, bits, mode ? kFDstBits + 1 : kFLitBits + 1, 15);
                if (!e) { flag = kSpanBad; break; }
            }
            const uint32_t l = e & 15u, xb = (e >> 6) & 15u, kind = e & kEntKind;
            if (mode ? kind != kEntLen : kind > kEntLen) {
                if (!mode && kind == kEntEob) { p += l; flag = kSpanEob; } else flag = kSpanBad;
                break;
            }
            const int val = (int)(e >> 16) + (int)((bits >> l) & ((1u << xb) - 1u));  // (l + xb <= 28 bits of the 32)
            p += l + xb;
            if (!mode && kind == kEntLen) { len = val; mode = 1; continue; }
            if (WRITE) {
                if (mode && val > op + n + extra) { flag = kSpanFar; break; }
                out[n] = mode ? (uint32_t)(len - 3) | ((uint32_t)val << 8) : (uint32_t)val;
            }
            n++;
            if (mode) { extra += len - 1; mode = 0; }
            if (p >= limit) break;
            if (MARKS) {
                const int wn = (int)((p - base) >> 5);
                if (wn != wcur) {  // the next symbol starts in a later word
                    marks[wcur * MS] = mark_pack(off, n - n0, extra - e0);
                    for (int w = wcur + 1; w < wn; w++) marks[w * MS] = kMarkNone;  // a symbol longer than a word
                    wcur = wn; off = (p - base) & 31u; n0 = n; e0 = extra;
                    fresh = true;
                }
            }
        }
        if (MARKS) { marks[wcur * MS] = mark_pack(off, n - n0, extra - e0); wcur++; }
    }
    if (MARKS) for (int w = wcur; w < kMarkWords; w++) marks[w * MS] = kMarkNone;
    if (flag == kSpanRun && p >= data_end) flag = kSpanBad;  // the data ended before the block did
    r.end = p; r.n = n; r.extra = extra; r.flag = flag;
}
template <bool WRITE>
CCT_HD void decode_span(const uint32_t *stage, uint32_t origin, uint32_t p, uint32_t limit, uint32_t data_end,
                        const FastTables &tb, uint32_t *out, int op, Span &r) {
    decode_core<WRITE, false, 1>(stage, origin, 0u, p, limit, data_end, tb, nullptr, false, out, op, r);
}
template <int MS>
CCT_HD void decode_span_marks(const uint32_t *stage, uint32_t origin, uint32_t base, uint32_t p, uint32_t limit,
                              uint32_t data_end, const FastTables &tb, uint32_t *marks, bool again, Span &r) {
    decode_core<false, true, MS>(stage, origin, base, p, limit, data_end, tb, marks, again, nullptr, 0, r);
}

// Block header (RFC 1951 3.2.3 - 3.2.7), single lane: leaves the code lengths in tb.lens (literal/length codes
// first, then distance codes) and returns the block type, or -1.  Stored: *nlen = LEN.  `br` stands after the header.
// With `lengths` off a dynamic block's code lengths are left to cl_entry / cl_walk: `br` stands at their first bit
// and the code-length code's table is at the start of the staging area.
CCT_HDN inline int inflate_block_header(BitReader &br, FastTables &tb, bool *last, int *nlen, int *ndist,
                                        bool lengths = true) {
    br.refill();
    *last = br.take(1) != 0;
    const int type = (int)br.take(2);
    if (type == 0) {
        br.drop(br.nbits & 7);
        br.refill();
        const uint32_t len = br.take(16), nl = br.take(16);
        if (br.bad || (len ^ 0xFFFFu) != nl) return -1;
        *nlen = (int)len;
        return 0;
    }
    if (type == 3) return -1;
    uint8_t *lens = tb.lens;
    if (type == 1) {
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        for (int i = 0; i < 30; i++) lens[288 + i] = 5;
        *nlen = 288; *ndist = 30;
        return 1;
    }
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const int nl = (int)br.take(5) + 257, nd = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
    if (nl > 286 || nd > 30) return -1;
    uint8_t cl[19];
    for (int i = 0; i < 19; i++) cl[i] = 0;
    for (int i = 0; i < ncode; i++) { br.refill(); cl[order[i]] = (uint8_t)br.take(3); }
    // the code-length code's little tables live in the (idle) staging area
    uint16_t *prim = reinterpret_cast<uint16_t *>(tb.stage), *cnt = prim + 128, *sy = cnt + 16;
    if (!build_decode_table(cl, 19, 7, prim, cnt, sy)) return -1;
    if (!lengths) { *nlen = nl; *ndist = nd; return br.bad ? -1 : 2; }
    int idx = 0;
    while (idx < nl + nd) {
        br.refill();
        const int sym = decode_symbol(br, prim, 7, cnt, sy);
        if (sym < 0 || br.bad) return -1;
        if (sym < 16) lens[idx++] = (uint8_t)sym;
        else {
            int rep, val = 0;
            if (sym == 16) { if (idx == 0) return -1; val = lens[idx - 1]; rep = 3 + (int)br.take(2); }
            else if (sym == 17) rep = 3 + (int)br.take(3);
            else rep = 11 + (int)br.take(7);
            if (idx + rep > nl + nd) return -1;
            while (rep--) lens[idx++] = (uint8_t)val;
        }
    }
    if (br.bad || lens[256] == 0) return -1;
    *nlen = nl; *ndist = nd;
    return 2;
}

// The code lengths of a dynamic block, in two steps.  The code-length symbols are at most 7 bits plus at most 7 extra
// bits and every one of them yields a length, so the nlen + ndist <= 316 lengths take fewer than kClBits bits.
// cl_entry   what a code-length symbol starting at a given bit would be (any thread, any bit): symbol | bits it takes
//            << 5 | value of its extra bits << 9, 0 for an unused code.  `bits` = 14 or more stream bits from there.
// cl_walk    follows the entries from bit 0 and writes the lengths (one thread).
constexpr int kClBits = 2240, kClWords = kClBits / 32 + 2;
constexpr int kClWordAt = 96, kClEntAt = 192;  // where the header's words and entries sit in the staging area (32-bit words)
static_assert(kClEntAt + kClBits / 2 <= kStageAlloc && kClWordAt + kClWords <= kClEntAt, "the staging area holds a header");
CCT_HD uint16_t cl_entry(const uint16_t *prim, uint32_t bits) {
    const uint32_t e = prim[bits & 127u];
    if (!e) return 0;
    const uint32_t l = e >> 12, sym = e & 0xFFFu;
    const uint32_t xb = sym < 16 ? 0u : sym == 16 ? 2u : sym == 17 ? 3u : 7u;
    return (uint16_t)(sym | ((l + xb) << 5) | (((bits >> l) & ((1u << xb) - 1u)) << 9));
}
CCT_HDN inline bool cl_walk(const uint16_t *ent, int total, uint8_t *lens, uint32_t *used) {
    uint32_t b = 0;
    int idx = 0;
    while (idx < total) {
        if (b >= (uint32_t)kClBits) return false;
        const uint32_t e = ent[b];
        if (!e) return false;
        const uint32_t sym = e & 31u;
        b += (e >> 5) & 15u;
        if (sym < 16) { lens[idx++] = (uint8_t)sym; continue; }
        int rep = (int)(e >> 9) + (sym == 18 ? 11 : 3), val = 0;
        if (sym == 16) { if (idx == 0) return false; val = lens[idx - 1]; }
        if (idx + rep > total) return false;
        while (rep--) lens[idx++] = (uint8_t)val;
    }
    *used = b;
    return lens[256] != 0;
}

// Pass 2 of inflate, sequential form (CPU harness): symbols -> bytes.
CCT_HDN inline void lz_expand(const uint32_t *syms, int ns, uint8_t *out) {
    int op = 0;
    for (int i = 0; i < ns; i++) {
        const uint32_t sy = syms[i];
        const int dist = (int)(sy >> 8);
        if (!dist) { out[op++] = (uint8_t)sy; continue; }
        const int len = (int)(sy & 255u) + 3;
        for (int k = 0; k < len; k++, op++) out[op] = out[op - dist];
    }
}

CCT_HD uint32_t adler32_bytes(const uint8_t *p, int n) {
    uint32_t a = 1, b = 0;
    while (n > 0) {
        int k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521u; b %= 65521u;
    }
    return (b << 16) | a;
}

// ------------------------------------------------------------------------------------------------
// CompaCT op stream (core.py:453-520 restated per SURVEY A.5)
// ------------------------------------------------------------------------------------------------
constexpr int32_t kSlotSecond = 0x40000000;  // slotB flag: second half of a meshed unit

// One op of the stream at payload[pos]: returns its byte length (1 or 2), sets *delta and *kind
// (0 pixel op, 1 JUMP with *delta = distance).  `end` = plen - 1: the reference's reader never consumes the
// EOF byte (core.py:137,142).  A FULL op whose low byte would be at/after `end` is reported as length 0.
CCT_HD int op_decode(const uint8_t *payload, int pos, int end, int *delta, int *kind) {
    const int b = payload[pos];
    *kind = 0;
    if ((b & 0xC0) == 0x80) { *kind = 1; *delta = b & 0x3F; return 1; }  // JUMP (core.py:484-494)
    if ((b & 0xF0) == 0xE0) {                                           // FULL (core.py:500-506)
        if (pos + 1 >= end) return 0;
        const int full = ((b << 8) | payload[pos + 1]) & 0xFFF;
        *delta = full > 2048 ? full - 4096 : full;                      // signed(x, 12): x > 2048 -> x - 4096
        return 2;
    }
    if ((b & 0x80) == 0) {                                              // DELTA (core.py:513-516)
        const int x = b & 0x7F;
        *delta = x > 64 ? x - 128 : x;
    } else *delta = 0;  // any other tag: the reference leaves the pixel unchanged
    return 1;
}

// The op stream is cut into fixed chunks so it can be decoded in parallel (op starts follow a
// 2-state automaton -- "this byte is the low byte of a FULL op that started in the previous chunk").
// Pass 1 summarises a chunk for one entry state; pass 2 (ops_link_chunks) chains the summaries of a stream;
// pass 3 (ops_emit_chunk) decodes the chunk again with its now-known entry state, op index and running value.
constexpr int kOpsChunk = 1024;
struct OpsChunkSum { int32_t nops[2]; int32_t dsum[2]; uint8_t exit_state[2]; uint8_t entry; uint8_t pad; int32_t op_base; int32_t val_base; };

CCT_HD void ops_scan_chunk(const uint8_t *payload, int lo, int hi, int end, int entry, int *nops, int *dsum, int *exit_state) {
    int pos = lo + entry, n = 0, sum = 0;
    const int stop = hi < end ? hi : end;
    while (pos < stop) {
        int delta, kind;
        const int len = op_decode(payload, pos, end, &delta, &kind);
        if (len == 0) { pos = stop; break; }
        pos += len;
        if (!kind) { n++; sum += delta; }
    }
    *nops = n; *dsum = sum; *exit_state = pos > hi ? 1 : 0;
}

// Sequential link over the chunk summaries of one stream.  Returns CCT_ST_* bits (truncated if fewer than n_px ops).
CCT_HDN inline uint32_t ops_link_chunks(OpsChunkSum *cs, int n_chunks, int n_px) {
    int state = 0, ops = 0, val = 0;
    for (int c = 0; c < n_chunks; c++) {
        cs[c].entry = (uint8_t)state; cs[c].op_base = ops; cs[c].val_base = val;
        ops += cs[c].nops[state]; val += cs[c].dsum[state];
        state = cs[c].exit_state[state];
    }
    return ops < n_px ? 64u : 0u;
}

// Pass 3: values of the chunk's pixel ops -> vals[k]; JUMP distances -> jflag[k / 16].  Returns status bits.
CCT_HD uint32_t ops_emit_chunk(const uint8_t *payload, int lo, int hi, int end, const OpsChunkSum &cs, int n_px,
                               uint16_t *vals, uint8_t *jflag) {
    int pos = lo + cs.entry, k = cs.op_base, value = cs.val_base;
    const int stop = hi < end ? hi : end;
    uint32_t st = 0;
    // values are written four at a time (one aligned 8-byte store): a thread owns a contiguous range of vals[],
    // and 2-byte stores from 32 threads would cost 32 memory transactions per value
    uint64_t quad = 0;
    while (pos < stop && k < n_px) {
        int delta, kind;
        const int len = op_decode(payload, pos, end, &delta, &kind);
        if (len == 0) break;
        pos += len;
        if (kind) {
            if ((k & 15) == 0) jflag[k >> 4] = (uint8_t)delta;
            else st |= 64u;  // a JUMP inside a unit is never produced by the encoder
            continue;
        }
        value += delta;
        if (value < 0 || value > 65535) st |= 32u;  // reference: OverflowError (core.py:506,516)
        const int q = k & 3;
        quad |= (uint64_t)(uint16_t)value << (16 * q);
        if (q == 3) {
            if (k - 3 >= cs.op_base) *reinterpret_cast<uint64_t *>(vals + k - 3) = quad;  // vals is 8-byte aligned
            else for (int j = cs.op_base; j <= k; j++) vals[j] = (uint16_t)(quad >> (16 * (j & 3)));
            quad = 0;
        }
        k++;
    }
    for (int j = (k & ~3) < cs.op_base ? cs.op_base : (k & ~3); j < k; j++) vals[j] = (uint16_t)(quad >> (16 * (j & 3)));
    return st;
}

// Whole-stream convenience wrapper (CPU harness): same three passes, sequentially.
CCT_HDN inline uint32_t ops_to_values(const uint8_t *payload, int plen, int n_px, uint16_t *vals, uint8_t *jflag,
                                      OpsChunkSum *cs) {
    const int end = plen - 1;
    const int n_chunks = (plen + kOpsChunk - 1) / kOpsChunk;
    for (int c = 0; c < n_chunks; c++)
        for (int e = 0; e < 2; e++) {
            int nops, dsum, ex;
            ops_scan_chunk(payload, c * kOpsChunk, (c + 1) * kOpsChunk, end, e, &nops, &dsum, &ex);
            cs[c].nops[e] = nops; cs[c].dsum[e] = dsum; cs[c].exit_state[e] = (uint8_t)ex;
        }
    uint32_t st = ops_link_chunks(cs, n_chunks, n_px);
    for (int c = 0; c < n_chunks; c++)
        st |= ops_emit_chunk(payload, c * kOpsChunk, (c + 1) * kOpsChunk, end, cs[c], n_px, vals, jflag);
    return st;
}

// JUMP flags -> unit table.  Slot q covers pixel ops [16q, 16q+16).  slotA[q] = block A of the unit,
// slotB[q] = -1 (plain unit) or partner block B (| kSlotSecond in the unit's second slot).
CCT_HDN inline uint32_t resolve_units(const uint8_t *jflag, int nb, int32_t *slotA, int32_t *slotB) {
    uint64_t win = 0;  // bit i: block cur+i already consumed as a partner
    int cur = 0, q = 0;
    while (q < nb) {
        while (win & 1ull) { win >>= 1; cur++; }
        int A = cur, j = jflag[q];
        if (j) {
            int B = A + j;
            if (B >= nb || q + 1 >= nb || ((win >> j) & 1ull)) return 64u;
            slotA[q] = A; slotB[q] = B;
            slotA[q + 1] = A; slotB[q + 1] = B | kSlotSecond;
            win |= 1ull << j;
            q += 2;
        } else {
            slotA[q] = A; slotB[q] = -1;
            q += 1;
        }
        win >>= 1; cur++;
    }
    return 0;
}

// raster index of pixel op k given the unit table and the traversal order
CCT_HD int32_t op_raster(int k, const int32_t *slotA, const int32_t *slotB, const int32_t *order) {
    int q = k >> 4, i = k & 15;
    int32_t A = slotA[q], Bf = slotB[q];
    if (Bf < 0) return order[A * kBlock + i];
    int m = i + ((Bf & kSlotSecond) ? 16 : 0);
    int32_t blk = (m & 1) ? (Bf & ~kSlotSecond) : A;
    return order[blk * kBlock + (m >> 1)];
}

}  // namespace cct
