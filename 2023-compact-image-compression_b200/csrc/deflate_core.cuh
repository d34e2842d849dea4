// deflate_core.cuh — the scalar building blocks of the bit-exact zlib-1.3 level-9 encoder, written as
// host/device functions so the CUDA kernels (cct_deflate.cu) and the CPU test harness
// (tests/hostsim) run the very same code.
//
// What is restated: the reference's entropy stage is `zlib.compress(data, level = 9)`
// (reference src/codec/core.py:340) = system libz 1.3 `deflate_slow` + `trees.c` (third-party, not in the
// reference tree; algorithm recorded in SURVEY.md Appendix B).  The decomposition used here:
//   1. hash chains are parse-independent (every position is inserted exactly once, in order): the chain of a
//      position is the list of earlier positions with its 3-byte hash, so with the positions SORTED by
//      (hash, position) (kernels k_hs_*) it is the piece of the sorted array just below the position;
//   2. `bucket_probe` runs longest_match from a cold start (best = 2) for EVERY position in parallel, scanning
//      that piece, and records (len, dist) for both chain budgets in a table;
//   3. zlib's lazy-evaluation state machine is walked over symbol starts: speculatively per window in
//      parallel (`walk_window`), then stitched per stream (`walk_stitch`);
//   4. `build_block` restates _tr_flush_block's tree construction and block-type choice;
//   5. the symbol bits are then packed in parallel from the code tables.
#pragma once
#include "cct_common.cuh"

// Work counters for the CPU harness (tests/hostsim defines CCT_STATS); compiled out of the kernels.
#if defined(CCT_STATS) && !defined(__CUDA_ARCH__)
extern long long cct_stat[8];
#define CCT_COUNT(i) (cct_stat[i]++)
#else
#define CCT_COUNT(i) ((void)0)
#endif

namespace cct {

// ------------------------------------------------------------------------------------------------
// symbol / table encodings
// ------------------------------------------------------------------------------------------------
// probe-table entry: bits 0..8 best length (0 = no match), bits 9..23 distance; in the budget-4096 table bit 30 says
// that the budget-1024 entry differs (and is stored in the second table; otherwise that table holds nothing)
constexpr uint32_t kProbeHiDiff = 0x40000000u;
constexpr uint32_t kProbeTail = 0x80000000u, kProbeTailLeader = 0x80000001u;  // (bucket_probe_quick: left to tail_collect / tail_entry)
CCT_HD uint32_t mtab_pack(int len, int dist) { return (uint32_t)len | ((uint32_t)dist << 9); }
CCT_HD int mtab_len(uint32_t e) { return (int)(e & 511u); }
CCT_HD int mtab_dist(uint32_t e) { return (int)((e >> 9) & 32767u); }

// LZ77 symbol: bits 0..7 literal byte or (len-3), bits 8..23 distance (0 = literal)
CCT_HD uint32_t sym_pack(int lc, int dist) { return (uint32_t)lc | ((uint32_t)dist << 8); }
CCT_HD int sym_lc(uint32_t s) { return (int)(s & 255u); }
CCT_HD int sym_dist(uint32_t s) { return (int)(s >> 8); }

CCT_HD uint32_t hash3(const uint8_t *p) {  // UPDATE_HASH x3 with hash_shift 5, 15-bit mask
    return (((uint32_t)p[0] << 10) ^ ((uint32_t)p[1] << 5) ^ (uint32_t)p[2]) & 0x7FFFu;
}

// length code / extra bits of lc = len-3 (trees.c _length_code, base_length, extra_lbits)
CCT_HD int len_code(int lc) {
    if (lc < 8) return lc;
    if (lc == 255) return 28;
    int m = ilog2((uint32_t)lc);
    return 4 * (m - 2) + 4 + ((lc >> (m - 2)) & 3);
}
CCT_HD int len_extra_bits(int code) { return (code < 8 || code == 28) ? 0 : (code >> 2) - 1; }
// distance code / extra bits of d = dist-1 (trees.c _dist_code, base_dist, extra_dbits)
CCT_HD int dist_code(int d) {
    if (d < 4) return d;
    int m = ilog2((uint32_t)d);
    return 2 * m + ((d >> (m - 1)) & 1);
}
CCT_HD int dist_extra_bits(int code) { return code < 4 ? 0 : (code >> 1) - 1; }
CCT_HD int bl_extra_bits(int code) { return code < 16 ? 0 : (code == 16 ? 2 : (code == 17 ? 3 : 7)); }

CCT_HD uint32_t bit_reverse(uint32_t code, int len) {  // trees.c bi_reverse: the low `len` bits of code, reversed
#if defined(__CUDA_ARCH__)
    return len ? __brev(code) >> (32 - len) : 0u;
#else
    uint32_t res = 0;
    for (int i = 0; i < len; i++) { res = (res << 1) | (code & 1u); code >>= 1; }
    return res;
#endif
}
// static trees (trees.c tr_static_init): literal/length lengths and bit-reversed codes
CCT_HD int static_llen(int n) { return n <= 143 ? 8 : (n <= 255 ? 9 : (n <= 279 ? 7 : 8)); }
CCT_HD uint32_t static_lcode(int n) {
    uint32_t c = n <= 143 ? 0x30u + (uint32_t)n
                          : (n <= 255 ? 0x190u + (uint32_t)(n - 144) : (n <= 279 ? (uint32_t)(n - 256) : 0xC0u + (uint32_t)(n - 280)));
    return bit_reverse(c, static_llen(n));
}

// ------------------------------------------------------------------------------------------------
// longest_match over hash-sorted positions
// ------------------------------------------------------------------------------------------------
struct LzView {
    const uint8_t *w;  // payload bytes
    int n;             // payload length
};

CCT_HD uint8_t load_byte(const uint8_t *p) {  // payload bytes are read-only for the whole deflate stage
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
CCT_HD int match_extend(const uint8_t *a, const uint8_t *b, int from, int cap) {
    int len = from;
    while (len < cap && load_byte(a + len) == load_byte(b + len)) len++;
    return len;
}

// A hash chain is nothing but the positions of one hash value in decreasing order, so with the positions of a
// stream SORTED by (hash, position) the chain of a target is the run of entries just below it in the sorted array:
// the pointer chase of longest_match becomes a scan of contiguous 8-byte entries (coalesced over the lanes of a warp,
// which hold neighbouring targets of the same bucket), and the links -- zlib's prev[] -- are never materialised.
// One entry per chain member (positions 1 .. n-3; position 0 doubles as NIL in zlib and is never a candidate):
//   a   position | kEntRun (the position starts a triple (c,c,c)) | kEntFirst (lowest position of its hash bucket)
//   q   w[p] | w[p+1] << 8 | w[p+2] << 16 | w[p+3] << 24   (0 past the end)
// One scan step handles either a single candidate or a whole RUN GROUP.  When the candidate `cur` starts with a triple
// (c,c,c), the chain entries below cur inside that run are exactly the consecutive positions cur-1, cur-2, ... down to
// the run start (all have the hash of (c,c,c)), hence consecutive entries of the sorted array.  If the target does not
// itself start with (c,c,c) -- the chain only shares the hash -- none of them matches 3 bytes.  If it starts with R >= 3
// copies of c, the match length of the i-th member is known in closed form from its count of leading c's, m + i:
// m+i if m+i < R, "R + whatever follows" if m+i == R, exactly R if m+i > R (R < cap), cap if R == cap.  Since only the
// first-visited candidate attaining the maximum matters, the group is consumed in one step and the scan
// jumps over its lower members; the chain budget and the distance limit are charged as if every member had been
// visited.  "Whatever follows" starts with the byte after either run: the payload is compared only when they agree.
// The run lengths around a position come from runs[p] = rl | rb << 16 (the bytes equal to w[p] immediately after /
// before p, both capped at 65535), read only on a run member.  (Carrying the counts in the entries of run members
// instead saves the walk a dependent load and measured 1 ms faster there, but cost the two sort passes 1.7 ms.)
// A single candidate is compared through the four payload bytes in its entry: one that merely shares the hash, or
// matches exactly three bytes, never touches the payload.
constexpr uint32_t kEntRun = 0x80000000u, kEntFirst = 0x40000000u, kEntPos = 0x3FFFFFFFu;
struct alignas(8) SortEnt { uint32_t a, q; };
CCT_HD int ent_pos(const SortEnt &e) { return (int)(e.a & kEntPos); }
CCT_HD bool quad_is_run(uint32_t q) { return ((q ^ (q >> 8)) & 0xFFFFu) == 0; }  // bytes 0, 1 and 2 are equal
CCT_HD uint32_t hash_quad(uint32_t q) { return (((q & 0xFFu) << 10) ^ ((q >> 3) & 0x1FE0u) ^ ((q >> 16) & 0xFFu)) & 0x7FFFu; }  // hash3 of bytes 0..2

// The sort key of a position: its 15-bit hash as (digit1 << 8) | digit2.  Any bijection of the hash would do -- only the
// positions of one bucket must end up together, in increasing order; this one folds the high bits into the low seven
// so that the first pass of the sort fills its 128 bins evenly.
CCT_HD uint32_t sort_digit1(uint32_t h) { return (h ^ (h >> 7)) & 0x7Fu; }
CCT_HD uint32_t sort_digit2(uint32_t h) { return h >> 7; }
CCT_HD uint32_t sort_key(uint32_t h) { return (sort_digit1(h) << 8) | sort_digit2(h); }

CCT_HD SortEnt load_ent(const SortEnt *p) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
    return SortEnt{v.x, v.y};
#else
    return *p;
#endif
}
// four payload bytes at any address (little-endian), as two aligned word loads on the device (payload rows are
// 16-byte aligned and padded, so the second word is inside the row whenever the first byte wanted from it is)
CCT_HD uint32_t load_u32_unaligned(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)a & 3u) * 8u;
    const uint32_t lo = __ldg(w);
    return sh ? __funnelshift_r(lo, __ldg(w + 1), sh) : lo;
#else
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
#endif
}
CCT_HD int count_trailing_zeros(uint32_t v) {  // v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
CCT_HD uint32_t load_u32(const uint32_t *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

struct BucketView {
    const uint8_t *w;       // payload bytes
    const SortEnt *ent;     // chain members sorted by (hash, position)
    const uint32_t *runs;   // per position: rl | rb << 16
    int n;                  // payload length
};
// the byte after the run a run member sits in (0 past the end or when the count is capped), as the group step needs it
CCT_HD uint32_t run_after(const BucketView &v, int p, int rl) {
    const int q = p + rl + 1;
    return (rl < 65535 && q < v.n) ? (uint32_t)load_byte(v.w + q) : 0u;
}

// Both walks below are written for a whole warp: every lane calls them together (a lane without a target passes
// active = false) and the candidate loops run while ANY lane is still walking, finished lanes idling.  The loops must
// be warp-uniform: with a plain per-thread loop the compiler gives the branches inside the body no reconvergence point
// before the loop exit, lanes that differ once never meet again, and the warp ends up stepping its lanes one after the
// other (measured: 6.5 of 32 threads per instruction).  On the host a "warp" is one thread.
#if defined(__CUDA_ARCH__)
#define CCT_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#else
#define CCT_WARP_ANY(p) (p)
#endif

// longest_match's candidate loop for the target at sorted index i, from a cold start (best_len 2), for both chain
// budgets: *lo for prev_length < good_length (4096 candidates), *hi for prev_length >= good_length (1024).  Same
// decisions, in the same order, as zlib's loop: a candidate replaces the best only when strictly longer, the
// distance limit ends the walk before the budget does, a match of the maximum useful length ends it at once.
// This is the general walk; bucket_probe_quick below settles most targets with far fewer instructions.
CCT_HD void bucket_probe(const BucketView &v, int i, const SortEnt &me, bool active, uint32_t *lo, uint32_t *hi) {
    *lo = *hi = 0;
    const SortEnt *pj = v.ent + i;  // the scan moves a pointer down the sorted array
    SortEnt e{0u, 0u};              // the candidate
    int s = 0, cap = 0, limit = 0, R = 0;
    uint32_t quad = 0, after = 0;
    if (active) {
        s = ent_pos(me);
        const int look = v.n - s;
        if (look < kMinMatch || (me.a & kEntFirst)) active = false;
        else {
            cap = look < kMaxMatch ? look : kMaxMatch;
            quad = me.q;
            if (me.a & kEntRun) {  // R = leading run of the target (at least one hash triple long), after = the byte after it
                const int rl = (int)(load_u32(v.runs + s) & 0xFFFFu);
                R = rl + 1 < cap ? rl + 1 : cap;
                after = run_after(v, s, rl);  // (meaningful when R < cap)
            }
            e = load_ent(--pj);
            if (s - ent_pos(e) > kMaxDist) active = false;  // deflate_slow's test on hash_head
            limit = s - kMaxDist;                           // every further candidate must lie above it
        }
    }
    int best = 2, bdist = 0, chain = kChainGood;
    bool second = false;  // the walk with the full budget continues the quartered one
    while (CCT_WARP_ANY(active)) {
        if (!active) continue;
        CCT_COUNT(2);
        int end = 0;  // how the walk ends: 0 not yet, 1 chain end / distance limit / maximum length, 2 chain budget
        const int cur = ent_pos(e);
        uint32_t bottom_a = e.a;  // flags of the entry whose predecessor the walk moves on to
        if (e.a & kEntRun) {
            // cur starts a triple (c,c,c): the chain entries below it inside the run form one group
            const uint32_t ri = load_u32(v.runs + cur);
            const int rl_g = (int)(ri & 0xFFFFu), m = rl_g + 1;  // leading c's of the top candidate
            int t = (int)(ri >> 16);                              // further candidates of this run below cur
            if (cur - t < 1) t = cur - 1;                         // position 0 is never in a chain
            const int count_dist = kMaxDist - (s - cur);
            int nvis = t + 1;
            if (chain < nvis) nvis = chain;
            if (count_dist < nvis) nvis = count_dist;
            if (R && (e.q & 0xFFu) == (quad & 0xFFu)) {
                const int i_key = m >= R ? 0 : R - m;  // first member with at least R leading c's
                if (R < cap && m > R) {
                    if (R > best) { best = R; bdist = s - cur; }
                } else if (i_key < nvis) {
                    const int c_al = cur - i_key;  // R < cap: exactly R leading c's, then the byte after the group's run
                    int L = cap;
                    if (R < cap) L = run_after(v, cur, rl_g) != after ? R : match_extend(v.w + c_al, v.w + s, R + 1, cap);
                    if (L > best) { best = L; bdist = s - c_al; if (L >= cap) end = 1; }
                } else {
                    const int L = m + nvis - 1;  // only members with fewer than R leading c's were reached
                    if (L > best) { best = L; bdist = s - (cur - (nvis - 1)); }
                }
            }  // else: the target does not start with (c,c,c) (a hash collision): no member matches 3 bytes
            if (!end && nvis < t + 1) end = nvis == count_dist ? 1 : 2;
            if (!end && t) { chain -= t; pj -= t; bottom_a = load_ent(pj).a; }
        } else if (((e.q ^ quad) & 0xFFFFFFu) == 0) {  // the hash triple matches
            int len = cap < kMinMatch ? cap : kMinMatch;
            if (e.q == quad && cap > kMinMatch) {
                // four bytes match: the payload decides (zlib's scan_end test is only a shortcut for len <= best)
                const uint8_t *a = v.w + cur, *b = v.w + s;
                len = (best > kMinMatch && load_byte(a + best) != load_byte(b + best)) ? 0 : match_extend(a, b, kMinMatch + 1, cap);
            }
            if (len > best) { best = len; bdist = s - cur; if (len >= cap) end = 1; }
        }
        if (!end) {
            // move on to the previous member of the bucket: the distance limit comes before the budget, as in zlib's
            // loop condition
            if (bottom_a & kEntFirst) end = 1;
            else {
                e = load_ent(--pj);
                chain--;
                if (ent_pos(e) <= limit) end = 1;
                else if (chain == 0) end = 2;
            }
        }
        if (!end) continue;
        const uint32_t r = best >= kMinMatch ? mtab_pack(best, bdist) : 0u;
        if (!second) *hi = r;
        *lo = r;
        if (end != 2 || second) { active = false; continue; }
        // The quartered budget ran out: the walk with the full budget visits the same candidates up to here, so it
        // simply goes on.  (A run group cut short by the budget left pj and chain untouched and is evaluated again with
        // the longer count -- its members seen so far are a prefix, so best/bdist come out as from a fresh walk.)
        second = true;
        chain += kChainFull - kChainGood;
    }
}

CCT_HD uint32_t run_class(int s, int n, uint32_t ri);  // (byte runs, settled once per run: further down)

// The same walk for the common targets, with far fewer instructions.  Returns false when the walk meets anything it
// does not handle (the caller then runs bucket_probe from the start); true with the table entry, valid for both
// budgets, otherwise (and for idle lanes).
//   * An ordinary position whose chain holds ordinary positions, none matching eight bytes or more, and ends within the
//     quartered budget: one 8-byte load and a few selects per candidate.  Most candidates DO start with the target's
//     three bytes (the hash is all but injective on small deltas) and are settled from the entry alone; the few that
//     match four bytes are only noted (a bit per candidate) and measured after every 32 candidates -- inside the loop
//     the two lanes that need the payload would stall the other thirty.  The result does not depend on that order: it
//     is the longest candidate, the nearest among equals.
//   * With `tails` set, the members of a byte run that run_collect settles once per run (run_class) are not walked at
//     all: the entry returned is the marker kProbeTail, or kProbeTailLeader for the start of the run.
//   * A position inside a byte run with at least the maximum match length still to go: its predecessor matches that
//     much, at distance 1, and nothing can beat it.  Likewise one with more than 4096 run members below it: they all
//     match exactly the rest of the run and use up the chain budget before anything else is seen.
CCT_HD bool bucket_probe_quick(const BucketView &v, int i, const SortEnt &me, bool active, bool tails, uint32_t *out) {
    *out = 0;
    // state of the lane's walk: 0 walking, 1 settled, 2 left to the general walk  (an integer, not two bools: the
    // compiler keeps bools that live across the loop as packed bytes and spends a third of the loop shuffling them)
    uint32_t state = 1;
    const SortEnt *pj = v.ent + i;
    SortEnt e{0u, 0u};
    int s = 0, limit = 0, best = 2, bdist = 0;
    uint32_t quad = 0;
    if (active) {
        s = ent_pos(me);
        quad = me.q;
        limit = s - kMaxDist;
        const int look = v.n - s;
        if (me.a & kEntRun) {
            state = 2;
            const uint32_t ri = load_u32(v.runs + s);
            const int cap = look < kMaxMatch ? look : kMaxMatch, R = (int)(ri & 0xFFFFu) + 1, rb = (int)(ri >> 16);
            // members of the target's own run below it (position 0 is never one): they are visited first, all give R
            const int own = rb < s - 1 ? rb : s - 1;
            const uint32_t cls = tails ? run_class(s, v.n, ri) : 0u;
            if (cls) state = 2u + cls;  // settled once per run (run_collect; the start announces the run even without a chain)
            else if (me.a & kEntFirst) state = 1;
            else if (own >= 1 && R >= cap) { best = cap; bdist = 1; state = 1; }
            else if (own > kChainFull) { best = R; bdist = 1; state = 1; }  // ... and exhaust either chain budget
        } else if (me.a & kEntFirst) state = 1;  // no chain: no match, whatever the budget
        else if (look < 8) state = 2;  // (the loop assumes a length cap >= 8)
        else {
            e = load_ent(--pj);
            state = s - ent_pos(e) > kMaxDist ? 1u : 0u;  // deflate_slow's test on hash_head
        }
    }
    // The candidates are visited in rounds of 32; a lane notes in `quads` (bit k = k-th candidate of the round) which
    // ones match four bytes.  After the round those are measured; kChainGood / 32 rounds exhaust the quartered budget.
    const SortEnt *round_top = pj;  // the first candidate of the round
    for (int round = 0; CCT_WARP_ANY(state == 0); round++) {
        if (round == kChainGood / 32) { if (state == 0) state = 2; break; }  // more than 1024 candidates: the budgets may differ
        uint32_t quads = 0;
        round_top = pj;
        for (uint32_t bit = 1; bit != 0 && CCT_WARP_ANY(state == 0); bit <<= 2) {
            if (state == 0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int half = 0; half < 2; half++) {  // two candidates per vote
                    CCT_COUNT(3);
                    const uint32_t ea = e.a, x = e.q ^ quad;
                    if ((x & 0xFFFFFFu) == 0 && best < kMinMatch) { best = kMinMatch; bdist = s - (int)(ea & kEntPos); }
                    if (x == 0) quads |= bit << half;
                    if (ea >= kEntFirst) { state = (ea & kEntRun) ? 2u : 1u; break; }  // a run member: the general walk takes over;
                    e = load_ent(--pj);                                                // the lowest position of the bucket: the chain ends
                    if (ent_pos(e) <= limit) { state = 1; break; }  // the distance limit
                }
            }
        }
        if (CCT_WARP_ANY(quads != 0 && state != 2)) {  // nearest first: a later candidate must be strictly longer
            const uint32_t quad2 = (quads != 0 && state != 2) ? load_u32_unaligned(v.w + s + 4) : 0u;
            while (CCT_WARP_ANY(quads != 0 && state != 2)) {
                if (quads != 0 && state != 2) {
                    const int k = count_trailing_zeros(quads);
                    quads &= quads - 1;
                    const int cur = ent_pos(load_ent(round_top - k));
                    const uint32_t y = load_u32_unaligned(v.w + cur + 4) ^ quad2;
                    if (y == 0) state = 2;  // eight bytes match: the general walk counts the rest
                    const int len = 4 + (count_trailing_zeros(y | 0x80000000u) >> 3);  // (< 8 <= cap: never the maximum length)
                    if (len > best) { best = len; bdist = s - cur; }
                }
            }
        }
    }
    if (state == 1 && best >= kMinMatch) *out = mtab_pack(best, bdist);
    if (state >= 3) *out = state == 4 ? kProbeTailLeader : kProbeTail;
    return state == 1 || state >= 3;
}

// ------------------------------------------------------------------------------------------------
// Byte runs, settled once per run
// ------------------------------------------------------------------------------------------------
// The chain members of a run X of byte c (start s0, `len` bytes, `end` the first position after it) are what the general
// walk spends most of its steps on -- every one of them steps through every earlier run group of the bucket of (c,c,c)
// -- and yet they all see the same thing.  A group G of the bucket is entered at its top member (exactly three c's
// ahead; member i below it has 3 + i); the member with as many leading c's as the target (R) matches R + A bytes, where
// A = the common prefix of what follows G and what follows X, those with more match R, those with fewer their own count.
//   * The START of X (R0 = min(len, 258) c's ahead, no member of X below it) is walked group by group like the general
//     walk does, but only once per run and with every group settled from its run counts.
//   * The members with 257 .. 3 c's still ahead and part of X below them (the "tail"; those with 258 or more are
//     settled by the quick walk) get (R, distance 1) from their own run first, so only a group's member with exactly R
//     leading c's -- end_G - R -- can beat that, and only if A >= 1.  For the target one position further R is one
//     less and that member lies one position further, too: its distance (end - end_G) and its place in the chain
//     (a difference of sorted indices) are THE SAME for every target of the tail; what varies is only whether the group
//     is long enough to hold such a member.  So the walk notes the few groups with A >= 1 that the chain budget and the
//     distance limit let the tail reach (dropping those a nearer record beats for every R), and each target reads its
//     entry off the records (tail_entry): the longest R + A, the nearest among equals -- which also reproduces the
//     early exit at the maximum length.
// Runs the walk cannot summarise (more records than it may keep, a group not entered at its top) go through the general
// walk, target by target.
//
// run_class: what the quick walk makes of a run member at s (ri = runs[s]): 0 = not settled here, 1 = a tail target,
// 2 = the start (it announces the run).  All members of a run agree: the run counts must not be capped, the run must
// not start at position 0 (never in a chain) and must end at least 255 bytes before the input does (the length cap is
// 258 for every target, and A can be measured without a bound check).
constexpr int kTailRecords = 8;
CCT_HD uint32_t run_class(int s, int n, uint32_t ri) {
    const int rl = (int)(ri & 0xFFFFu), rb = (int)(ri >> 16);
    if (rl < kMinMatch - 1 || rb + rl + 1 > 65535 || s - rb < 1 || n - (s + rl + 1) < kMaxMatch - kMinMatch) return 0;
    if (rb == 0) return 2;
    return rl + 1 < kMaxMatch ? 1u : 0u;
}

struct TailGroup {  // (8 bytes: the records of 32 runs sit side by side in shared memory)
    uint32_t a, b;  // a = rank | max_r << 16, b = dist | extra << 16
};
// rank: place in every tail target's chain of the group's member with exactly R leading c's (1 = nearest candidate);
// dist: its distance; max_r: the group holds such a member for R <= max_r; extra: A, 1 .. 255
CCT_HD TailGroup tail_group(int rank, int dist, int max_r, int extra) {
    return TailGroup{(uint32_t)rank | ((uint32_t)max_r << 16), (uint32_t)dist | ((uint32_t)extra << 16)};
}

// One walk over the run groups below the run whose start is entry i0 of the sorted array, warp-wide like the walks above
// (a lane per run: neighbouring runs of a bucket see the same groups).  Writes the start's table entries (both budgets)
// and the tail's records out[k * stride]; returns their number, -1 when the run must be left to the general walk.
CCT_HDN inline int run_collect(const BucketView &v, int i0, bool active, TailGroup *out, int stride, uint32_t *start_lo,
                               uint32_t *start_hi) {
    *start_lo = *start_hi = 0;
    const SortEnt *pj = v.ent + i0;
    int s0 = 0, len = 0, end = 0, r0 = 0, rank = 0, cnt = 0, limit = 0;
    uint32_t after = 0, bottom_a = 0;
    int best[2] = {2, 2}, bdist[2] = {0, 0};  // the start's walk, budget kChainGood / kChainFull
    bool live[2] = {false, false};
    SortEnt next{0u, 0u};  // the entry below *pj when it has been fetched along with *pj (below a group)
    bool have_next = false;
    if (active) {
        const SortEnt first = load_ent(pj);
        s0 = ent_pos(first);
        len = (int)(load_u32(v.runs + s0) & 0xFFFFu) + 1;
        end = s0 + len;
        r0 = len < kMaxMatch ? len : kMaxMatch;
        after = load_byte(v.w + end);
        bottom_a = first.a;
        limit = s0 - kMaxDist;
        live[0] = live[1] = true;
    }
    while (CCT_WARP_ANY(active)) {
        if (!active) continue;
        if (bottom_a & kEntFirst) { active = false; continue; }
        --pj;
        const SortEnt e = have_next ? next : load_ent(pj);
        have_next = false;
        rank++;  // e is the rank-th candidate of the start
        const int cur = ent_pos(e);
        // the start's loop condition (the first candidate may lie one further than the others); the tail's reach ends earlier
        if (rank > kChainFull || (rank == 1 ? s0 - cur > kMaxDist : cur <= limit)) { active = false; continue; }
        if (rank > kChainGood) live[0] = false;
        bottom_a = e.a;
        if (!(e.a & kEntRun)) continue;  // shares the hash only: one step of the chain
        const uint32_t ri = load_u32(v.runs + cur);
        if ((ri & 0xFFFFu) != (uint32_t)(kMinMatch - 1)) { cnt = -1; active = false; continue; }  // (a group is entered at its top member)
        int t = (int)(ri >> 16);
        if (t == 65535) { cnt = -1; active = false; continue; }
        if (cur - t < 1) t = cur - 1;
        const int end_g = cur + kMinMatch;
        const int extra = load_byte(v.w + end_g) == after ? match_extend(v.w + end_g, v.w + end, 1, kMaxMatch - kMinMatch) : 0;
        // the start: members 0 .. nvis-1 of the group are visited
        const int count_dist = kMaxDist - (s0 - cur), i_key = r0 - kMinMatch;
        bool more = false;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < 2; b++) {
            if (!live[b]) continue;
            int nvis = (b ? kChainFull : kChainGood) - rank + 1;
            if (t + 1 < nvis) nvis = t + 1;
            if (count_dist < nvis) nvis = count_dist;
            if (i_key < nvis) {
                int L = r0 + extra;
                if (L > kMaxMatch || r0 == kMaxMatch) L = kMaxMatch;
                if (L > best[b]) { best[b] = L; bdist[b] = s0 - (cur - i_key); if (L >= kMaxMatch) live[b] = false; }
            } else {
                const int L = kMinMatch + nvis - 1;
                if (L > best[b]) { best[b] = L; bdist[b] = s0 - (cur - (nvis - 1)); }
            }
            if (nvis < t + 1) live[b] = false;
            more |= live[b];
        }
        // the tail: its targets meet the member with their R at rank + len - 3, end - end_g away
        const int rank_al = rank + len - kMinMatch, dist = end - end_g;
        const bool tail_reach = len > kMinMatch && rank_al <= kChainFull && dist < kMaxDist;
        if (tail_reach && extra > 0) {
            const int max_r = kMinMatch + t < kMaxMatch ? kMinMatch + t : kMaxMatch;
            bool beaten = false;  // by a nearer record that is at least as long for every R
            for (int k = 0; k < cnt; k++) {
                const TailGroup g = out[k * stride];
                if ((int)(g.b >> 16) >= extra && (int)(g.a >> 16) >= max_r) beaten = true;
            }
            if (!beaten) {
                if (cnt == kTailRecords) { cnt = -1; active = false; continue; }
                out[cnt++ * stride] = tail_group(rank_al, dist, max_r, extra);
            }
        }
        if (!more && !tail_reach) { active = false; continue; }
        // on to the group's lowest member, and -- in the same round trip -- the entry below it
        pj -= t; rank += t;
        have_next = pj > v.ent;
        if (have_next) next = load_ent(pj - 1);
        if (t) bottom_a = load_ent(pj).a;
    }
    if (cnt >= 0) {
        *start_hi = best[0] >= kMinMatch ? mtab_pack(best[0], bdist[0]) : 0u;
        *start_lo = best[1] >= kMinMatch ? mtab_pack(best[1], bdist[1]) : 0u;
    }
    return cnt;
}

// Table entries (both budgets) of the tail target with R leading c's.
CCT_HD void tail_entry(const TailGroup *g, int stride, int cnt, int R, uint32_t *lo, uint32_t *hi) {
    int blo = R, dlo = 1, bhi = R, dhi = 1;
    for (int k = 0; k < cnt; k++) {
        const TailGroup r = g[k * stride];
        if (R > (int)(r.a >> 16)) continue;
        int L = R + (int)(r.b >> 16);
        if (L > kMaxMatch) L = kMaxMatch;
        if (L > blo) { blo = L; dlo = (int)(r.b & 0xFFFFu); }  // (every record is within the full budget)
        if ((int)(r.a & 0xFFFFu) <= kChainGood && L > bhi) { bhi = L; dhi = (int)(r.b & 0xFFFFu); }
    }
    *lo = mtab_pack(blo, dlo);
    *hi = mtab_pack(bhi, dhi);
}

// ------------------------------------------------------------------------------------------------
// deflate_slow as a walk over symbol starts
// ------------------------------------------------------------------------------------------------
// deflate_slow (deflate.c) emits the literal for byte s-1 while standing at s; re-indexed by the position a
// symbol STARTS at, its lazy state machine reads:
//     at symbol start p, with either no pending match ("cold") or the match (pl, pd) found at p by the
//     previous step's look-ahead:
//         cur = pending, or the cold search at p (prev_length 2, TOO_FAR rule)
//         no cur                      -> LITERAL(w[p]);  next start p+1, cold
//         search at p+1 beats cur     -> LITERAL(w[p]);  next start p+1, pending = that longer match
//         otherwise                   -> MATCH(cur);     next start p+len, cold
// A search result comes from the probe table: it does not depend on prev_length beyond "must be longer" and
// the chain budget (one table per budget; they differ only where a chain holds more than 1024 candidates), so a
// pending match at p is simply the table entry of p and the walk state is a function of p alone wherever the
// two tables agree.  Because the state is (almost always) a
// function of p, walks started at different positions coincide from the first symbol start they share: the
// parse is done speculatively per window in parallel (walk_window), then stitched by one sequential pass per
// stream that only re-walks up to the point where the true path joins the speculative one (walk_stitch).
struct Walk {
    int p, pl, pd;  // symbol start, pending match length (0 = cold) and distance
};

struct ProbeTab {
    const uint32_t *lo, *hi;  // entries for chain budget 4096 (prev_length < 32) and 1024 (prev_length >= 32)
};

CCT_HD uint32_t probe_lo(const ProbeTab &mt, int p) { return mt.lo[p] & ~kProbeHiDiff; }
CCT_HD uint32_t probe_hi(const ProbeTab &mt, int p) { const uint32_t e = mt.lo[p]; return (e & kProbeHiDiff) ? mt.hi[p] : e; }

CCT_HD bool lookup_match(const LzView &v, const ProbeTab &mt, int p, int prev_len, int *len, int *dist) {
    const int look = v.n - p;
    if (look < kMinMatch) return false;
    const uint32_t e = prev_len >= kGoodLen ? probe_hi(mt, p) : probe_lo(mt, p);
    if (e == 0 || prev_len >= kMaxMatch || prev_len >= look) return false;
    int ml = 0, md = 0;
    if (mtab_len(e) > prev_len) { ml = mtab_len(e); md = mtab_dist(e); }
    if (ml == kMinMatch && md > kTooFar) ml = 0;
    *len = ml; *dist = md;
    return ml != 0;
}

// One symbol: returns it and advances the walk.
CCT_HD uint32_t walk_step(const LzView &v, const ProbeTab &mtab, Walk &w) {
    const int p = w.p;
    int l0 = w.pl, d0 = w.pd;
    if (!l0 && !lookup_match(v, mtab, p, 2, &l0, &d0)) { w.p = p + 1; return sym_pack(v.w[p], 0); }
    int l1, d1;
    if (lookup_match(v, mtab, p + 1, l0, &l1, &d1)) { w.p = p + 1; w.pl = l1; w.pd = d1; return sym_pack(v.w[p], 0); }
    w.p = p + l0; w.pl = 0; w.pd = 0;
    return sym_pack(l0 - kMinMatch, d0);
}

// Do two walks standing on the same symbol start continue identically?
CCT_HD bool walk_same_state(const ProbeTab &mt, const Walk &a, const Walk &b) {
    if (a.pl == b.pl && a.pd == b.pd) return true;
    return !(mt.lo[a.p] & kProbeHiDiff);  // one entry for both budgets: the state is a function of the position
}

// The step as a function of the position alone.  From a cold start at p the machine needs lookup(p, 2) and, if that
// finds a match of length l0, lookup(p + 1, l0); when both come straight from complete table entries the symbol and
// the next start are known without walking -- and a walk that arrives at p+1 with the pending match this step found
// continues exactly like a cold one there, provided the pending entry is the one a cold lookup reads (always the case
// below good_length; above it the two budget tables must agree).  walk_precompute evaluates this for every position
// in parallel (k_next): sym[p] = the symbol a walk standing at p emits (the advance follows from it), or kSymSlow
// when the step depends on how the walk got there.  The walks themselves (walk_window /
// walk_stitch) then chase one 4-byte entry per symbol and fall back to walk_step only on kSymSlow entries, whose
// symbol they record in place with the marker kept.
constexpr uint32_t kSymSlow = 0x80000000u;
CCT_HD int sym_advance(uint32_t sym) { return sym_dist(sym & ~kSymSlow) ? sym_lc(sym) + kMinMatch : 1; }

CCT_HD uint32_t walk_precompute(const LzView &v, const ProbeTab &mt, int p) {
    const uint32_t lit = sym_pack(v.w[p], 0);
    const int look = v.n - p;
    if (look < kMinMatch) return lit;
    const uint32_t e = probe_lo(mt, p);  // lookup(p, 2)
    if (e == 0) return lit;
    const int l0 = mtab_len(e), d0 = mtab_dist(e);
    if (l0 == kMinMatch && d0 > kTooFar) return lit;
    const uint32_t match = sym_pack(l0 - kMinMatch, d0);
    const int look1 = look - 1;  // lookup(p + 1, l0)
    if (look1 < kMinMatch) return match;
    const uint32_t e1 = l0 >= kGoodLen ? probe_hi(mt, p + 1) : probe_lo(mt, p + 1);
    if (e1 == 0 || l0 >= kMaxMatch || l0 >= look1) return match;
    if (mtab_len(e1) <= l0) return match;  // (a longer match is at least 4 long: the TOO_FAR rule cannot apply)
    if (l0 >= kGoodLen && probe_lo(mt, p + 1) != e1) return kSymSlow;  // the pending match is not what a cold lookup reads
    return lit;
}

// One symbol of a walk over the precomputed table.  `record`: store the symbol of a slow step (the true walk and the
// speculative window walks do; the stitch's re-run of a speculative walk does not).
CCT_HD void walk_step_tab(const LzView &v, const ProbeTab &mtab, uint32_t *sym, Walk &w, bool record) {
    const int q = w.p;
    if (w.pl == 0) {
        const uint32_t s = sym[q];
        if (!(s & kSymSlow)) { w.p = q + sym_advance(s); return; }
    }
    const uint32_t s = walk_step(v, mtab, w);
    if (record) sym[q] = s | kSymSlow;
}

// Symbol starts are kept as a bit mask, one 32-bit word per 32 positions.
CCT_HD void mark_set(uint32_t *mask, int p) { mask[p >> 5] |= 1u << (p & 31); }
CCT_HD void mark_clear(uint32_t *mask, int p) { mask[p >> 5] &= ~(1u << (p & 31)); }

// Speculative walk of window [ws, we) (ws a multiple of 32): assumes a cold symbol start at ws.  Writes the mask words
// of the window (bits of the symbol starts it visits, zeros elsewhere) and returns the state it leaves the window with.
CCT_HD Walk walk_window(const LzView &v, const ProbeTab &mtab, int ws, int we, uint32_t *mask, uint32_t *sym) {
    Walk w{ws, 0, 0};
    int word = ws >> 5;
    uint32_t bits = 0;
    while (w.p < we) {
        const int q = w.p;
        walk_step_tab(v, mtab, sym, w, true);
        while (word < (q >> 5)) { mask[word++] = bits; bits = 0; }
        bits |= 1u << (q & 31);
    }
    const int last = (we + 31) >> 5;
    while (word < last) { mask[word++] = bits; bits = 0; }
    return w;
}

// Sequential stitch of one stream: window after window, advance the true walk T and (re-running it) the
// speculative walk S of the window in position order; S's marks that T does not confirm are cleared, T's
// symbols are recorded, and as soon as both stand on the same symbol start in the same state the rest of the
// window's speculation is accepted wholesale (T jumps to the recorded exit state).
CCT_HDN inline void walk_stitch(const LzView &v, const ProbeTab &mtab, int win, uint32_t *mask, uint32_t *sym,
                                const Walk *exits) {
    Walk T{0, 0, 0};
    const int n = v.n;
    for (int k = 0, ws = 0; ws < n; k++, ws += win) {
        const int we = ws + win < n ? ws + win : n;
        if (T.p == ws && T.pl == 0) { T = exits[k]; continue; }  // speculation started from the true state
        Walk S{ws, 0, 0};
        for (;;) {
            const bool t_in = T.p < we, s_in = S.p < we;
            if (!t_in && !s_in) break;
            if (t_in && s_in && T.p == S.p) {
                if (walk_same_state(mtab, T, S)) { T = exits[k]; break; }
                // same start, different pending match over an incomplete probe: both move on
                walk_step_tab(v, mtab, sym, S, false);  // (first: T's step may rewrite the entry S must still see as slow)
                walk_step_tab(v, mtab, sym, T, true);
            } else if (s_in && (!t_in || S.p < T.p)) {
                mark_clear(mask, S.p);  // not on the true path
                walk_step_tab(v, mtab, sym, S, false);
            } else {
                mark_set(mask, T.p);
                walk_step_tab(v, mtab, sym, T, true);
            }
        }
    }
}

// Block boundaries.  zlib flushes a block as soon as 16383 symbols have been tallied inside its
// loop; the literal for the last byte is tallied after the loop and never triggers a flush.  With symbols
// numbered in stream order, block j < nflush ends where symbol 16383 (j+1) starts (recorded by the compaction
// pass), or at n when the input ends exactly there (then zlib appends an empty last block).
CCT_HD int finish_blocks(int nsym, uint32_t last_sym, int n, int32_t *blk_end) {
    const int lastlit = (nsym > 0 && sym_dist(last_sym) == 0) ? 1 : 0;
    const int nf = (nsym - lastlit) / kSymEnd;
    if (nf > 0 && nf * kSymEnd == nsym) blk_end[nf - 1] = n;
    return nf;
}

// ------------------------------------------------------------------------------------------------
// trees.c: build_tree / gen_bitlen / gen_codes / scan_tree / send_tree / block-type choice
// ------------------------------------------------------------------------------------------------
constexpr int kLCodes = 286, kDCodes = 30, kBLCodes = 19, kHeapSize = 2 * kLCodes + 1, kMaxBits = 15;
constexpr int kHdrWords = 160;  // dynamic-block header bit string capacity (worst case 4498 bits)

// One Huffman tree under construction (ct_data with Freq/Len/Code split out).  The kernel's speed is the number of
// blocks resident per SM, so the arrays are as small as they can be: frequencies of the leaves only (an inner node's
// lives in its queue entry), lengths in bytes, the codes in the place of the frequencies (dead once the lengths and
// the cost sums are known), and no Dad links at all: inner nodes are numbered in the order they are made, so the
// k-th pair taken off the queue has the parent elems + k, and the order array says which pair a node was in.
struct Tree {
    uint16_t freq[kLCodes + 2];  // leaves only; after gen_bitlen: their codes
    uint8_t len[kHeapSize + 3];
    int max_code;
    CCT_HD uint16_t *code() { return freq; }
    CCT_HD const uint16_t *code() const { return freq; }
};

struct SmallTree {  // distance (30 symbols) and bit-length (19 symbols) trees
    uint16_t freq[kDCodes + 2];
    uint8_t len[2 * kDCodes + 3];
    int max_code;
    CCT_HD uint16_t *code() { return freq; }
    CCT_HD const uint16_t *code() const { return freq; }
};

struct TreeScratch {
    // The priority queue of build_tree, one word per entry: Freq << 16 | depth << 10 | node.  trees.c's
    // smaller(n, m) = Freq[n] < Freq[m] || (equal && depth[n] <= depth[m]) is exactly (q[n] >> 10) <= (q[m] >> 10);
    // with the node in the same word a sift level is one 8-byte load of both children instead of a chain of
    // heap[] -> Freq[] -> depth[] loads.  (Freq <= 16384 + 2 symbols per block, depth <= 22 < 64, node < 573.)
    // (At most one entry per leaf is ever queued: a merge takes two out and puts one in.)
    alignas(8) uint32_t pq[kLCodes + 6];
    int16_t heap[kHeapSize];  // nodes in the order they left the queue (zlib keeps them in the top of its heap array)
    uint16_t bl_count[kMaxBits + 1];
    int heap_len, heap_max;
    int64_t opt_len, static_len;
};

#if !defined(__CUDACC__)
struct uint2 { uint32_t x, y; };  // host build of the test harness
#endif
constexpr uint32_t kPqNone = 0xFFFFFFFFu;  // sentinel after the last entry: never "smaller" than a real one

CCT_HDN inline void pq_down(TreeScratch &s, int k, const int heap_len) {  // (heap_len by value: s.pq stores would force a reload)
    // (a >> 10) <= (b >> 10) is a <= (b | 1023): one OR per comparison instead of two shifts
    const uint32_t v = s.pq[k];
    int j = k << 1;
    while (j <= heap_len) {
        // children j (even: 8-byte aligned) and j + 1 (the sentinel when j is the last entry)
        const uint2 ch = *reinterpret_cast<const uint2 *>(&s.pq[j]);
        uint32_t c = ch.x;
        if (ch.y <= (c | 1023u)) { j++; c = ch.y; }
        if (v <= (c | 1023u)) break;
        s.pq[k] = c;
        k = j; j <<= 1;
    }
    s.pq[k] = v;
}

// kind: 0 literal/length tree, 1 distance tree, 2 bit-length tree
template <class T>
CCT_HDN inline void build_tree(TreeScratch &s, T &t, int kind) {
    const int elems = kind == 0 ? kLCodes : (kind == 1 ? kDCodes : kBLCodes);
    const int base = kind == 0 ? 257 : 0;
    const int max_length = kind == 2 ? 7 : kMaxBits;
    const int heap_size = 2 * elems + 1;  // zlib uses one HEAP_SIZE array; only relative order matters
    int n, m, max_code = -1, node;
    // the queue length, the stack cursor and the two cost sums live in registers (the struct sits in shared memory)
    int heap_len = 0, heap_max = heap_size;
    int64_t opt_len = s.opt_len, static_len = s.static_len;
    for (n = 0; n < elems; n++) {
        if (t.freq[n] != 0) { s.pq[++heap_len] = ((uint32_t)t.freq[n] << 16) | (uint32_t)(max_code = n); }
        else t.len[n] = 0;
    }
    while (heap_len < 2) {  // force at least two codes
        node = max_code < 2 ? ++max_code : 0;
        s.pq[++heap_len] = (1u << 16) | (uint32_t)node;
        t.freq[node] = 1;
        opt_len--;
        if (kind == 0) static_len -= static_llen(node);
        else if (kind == 1) static_len -= 5;
    }
    t.max_code = max_code;
    s.pq[heap_len + 1] = kPqNone;
    for (n = heap_len / 2; n >= 1; n--) pq_down(s, n, heap_len);
    node = elems;
    do {
        const uint32_t nq = s.pq[1];  // pqremove
        s.pq[1] = s.pq[heap_len]; s.pq[heap_len] = kPqNone; heap_len--;
        pq_down(s, 1, heap_len);
        const uint32_t mq = s.pq[1];
        n = (int)(nq & 1023u); m = (int)(mq & 1023u);
        s.heap[--heap_max] = (int16_t)n; s.heap[--heap_max] = (int16_t)m;
        const uint32_t f = (nq >> 16) + (mq >> 16);
        const uint32_t dn = (nq >> 10) & 63u, dm = (mq >> 10) & 63u;
        s.pq[1] = (f << 16) | ((((dn >= dm ? dn : dm) + 1u) & 63u) << 10) | (uint32_t)node;
        node++;
        pq_down(s, 1, heap_len);
    } while (heap_len >= 2);
    s.heap[--heap_max] = (int16_t)(s.pq[1] & 1023u);

    // gen_bitlen
    int h, bits, overflow = 0;
    for (bits = 0; bits <= kMaxBits; bits++) s.bl_count[bits] = 0;
    t.len[s.heap[heap_max]] = 0;
    for (h = heap_max + 1; h < heap_size; h++) {
        n = s.heap[h];
        bits = t.len[elems + ((heap_size - 1 - h) >> 1)] + 1;  // the parent of the pair n left the queue in
        if (bits > max_length) { bits = max_length; overflow++; }
        t.len[n] = (uint8_t)bits;
        if (n > max_code) continue;  // internal node
        s.bl_count[bits]++;
        int xbits = 0;
        if (n >= base) xbits = kind == 0 ? len_extra_bits(n - base) : (kind == 1 ? dist_extra_bits(n) : bl_extra_bits(n));
        const int f = t.freq[n];
        opt_len += f * (bits + xbits);
        if (kind == 0) static_len += f * (static_llen(n) + xbits);
        else if (kind == 1) static_len += f * (5 + xbits);
    }
    if (overflow) {
        do {
            bits = max_length - 1;
            while (s.bl_count[bits] == 0) bits--;
            s.bl_count[bits]--; s.bl_count[bits + 1] += 2; s.bl_count[max_length]--;
            overflow -= 2;
        } while (overflow > 0);
        for (bits = max_length; bits != 0; bits--) {
            n = s.bl_count[bits];
            while (n != 0) {
                m = s.heap[--h];
                if (m > max_code) continue;
                if (t.len[m] != (unsigned)bits) {
                    opt_len += ((int64_t)bits - t.len[m]) * t.freq[m];
                    t.len[m] = (uint8_t)bits;
                }
                n--;
            }
        }
    }
    s.heap_len = heap_len; s.heap_max = heap_max; s.opt_len = opt_len; s.static_len = static_len;
    // gen_codes
    uint32_t next_code[kMaxBits + 1], code = 0;
    next_code[0] = 0;
    for (bits = 1; bits <= kMaxBits; bits++) { code = (code + s.bl_count[bits - 1]) << 1; next_code[bits] = code & 0xFFFFu; }
    for (n = 0; n <= max_code; n++) {
        int len = t.len[n];
        if (len == 0) continue;
        t.code()[n] = (uint16_t)bit_reverse(next_code[len]++, len);  // (in place of freq[n])
    }
}

// tiny LSB-first bit writer into a uint32 word array (dynamic-block header string)
struct HdrBits {
    uint32_t *w; int nbits;
    CCT_HD void put(uint32_t value, int length) {
        int word = nbits >> 5, off = nbits & 31;
        w[word] |= value << off;
        if (off + length > 32) w[word + 1] |= value >> (32 - off);
        nbits += length;
    }
};

// scan_tree (count into blfreq) when bl == nullptr, send_tree (emit) otherwise — one control flow
template <class T>
CCT_HDN inline void tree_rle(const T &t, int max_code, uint16_t *blfreq, const SmallTree *bl, HdrBits *out) {
    int prevlen = -1, curlen, nextlen = t.len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    for (int n = 0; n <= max_code; n++) {
        curlen = nextlen;
        nextlen = (n + 1 <= max_code) ? t.len[n + 1] : 0xFFFF;
        if (++count < max_count && curlen == nextlen) continue;
        else if (count < min_count) {
            if (blfreq) blfreq[curlen] += (uint16_t)count;
            else do { out->put(bl->code()[curlen], bl->len[curlen]); } while (--count != 0);
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (blfreq) blfreq[curlen]++;
                else { out->put(bl->code()[curlen], bl->len[curlen]); count--; }
            }
            if (blfreq) blfreq[16]++;
            else { out->put(bl->code()[16], bl->len[16]); out->put((uint32_t)(count - 3), 2); }
        } else if (count <= 10) {
            if (blfreq) blfreq[17]++;
            else { out->put(bl->code()[17], bl->len[17]); out->put((uint32_t)(count - 3), 3); }
        } else {
            if (blfreq) blfreq[18]++;
            else { out->put(bl->code()[18], bl->len[18]); out->put((uint32_t)(count - 11), 7); }
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

// Result of _tr_flush_block's decisions for one block; consumed by the bit-packing kernel.
struct BlockDesc {
    uint32_t sym_begin, nsym;        // symbol range inside the stream's symbol array
    uint32_t pos_begin, stored_len;  // payload bytes covered (for stored blocks)
    uint32_t type, last;             // 0 stored, 1 static, 2 dynamic
    uint32_t hdr_bits;               // header bit count incl. the 3 block-header bits (dynamic: + trees)
    uint32_t body_bits;              // symbol bits + EOB (0 for stored)
    uint64_t bit_start;              // absolute bit offset in the zlib stream (filled by the layout pass)
    uint16_t lcode[kLCodes + 2]; uint8_t llen[kLCodes + 2];
    uint16_t dcode[kDCodes + 2]; uint8_t dlen[kDCodes + 2];
    uint32_t hdr[kHdrWords];
};

struct BlockWork {  // scratch for build_block (shared or local memory)
    Tree lt; SmallTree dt, bt; TreeScratch sc;
};

// Start of zlib's 64 KiB window when deflate_slow's loop stands at position t (deflate.c fill_window): the window slides
// by 32 KiB at the first loop top where fewer than MIN_LOOKAHEAD (262) bytes are left in it -- or, once the input is
// exhausted, at every loop top -- and strstart has reached wsize + MAX_DIST (65274) inside it.  A block whose first byte
// lies before the window (block_start < 0) cannot be emitted as a stored block (_tr_flush_block: buf == NULL).  The loop
// top of the iteration that flushes a block is the start of the block's last symbol + 1 (the end of input for the last
// block).
CCT_HD int64_t window_base_at(int64_t t, int64_t n) {
    int64_t k = t >= 65275 ? (t - 65275) / 32768 + 1 : 0;  // slides that have certainly happened
    for (;; k++) {
        const int64_t base = 32768 * k;
        const int64_t thr = n >= base + 65536 ? base + 65275 : (base + 65274 > n - 261 ? base + 65274 : n - 261);
        if (t < thr) return base;
    }
}

// lt.freq / dt.freq must hold the block's symbol frequencies (incl. freq[256] = 1) on entry, every
// other field zero; can_store = the block's first byte is still inside zlib's window when it is flushed.  Fills
// desc->type/last/stored_len/hdr/hdr_bits/body_bits and the code tables.  `desc` is
// written only after the three trees are built, so it may occupy the storage of wk.sc.pq / wk.sc.heap.
CCT_HDN inline void build_block(BlockWork &wk, BlockDesc *desc, uint32_t stored_len, uint32_t last, bool can_store) {
    const uint8_t bl_order[kBLCodes] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    wk.sc.opt_len = 0; wk.sc.static_len = 0;
    build_tree(wk.sc, wk.lt, 0);
    build_tree(wk.sc, wk.dt, 1);
    tree_rle(wk.lt, wk.lt.max_code, wk.bt.freq, (const SmallTree *)nullptr, (HdrBits *)nullptr);
    tree_rle(wk.dt, wk.dt.max_code, wk.bt.freq, (const SmallTree *)nullptr, (HdrBits *)nullptr);
    build_tree(wk.sc, wk.bt, 2);
    int max_blindex;
    for (max_blindex = kBLCodes - 1; max_blindex >= 3; max_blindex--)
        if (wk.bt.len[bl_order[max_blindex]] != 0) break;
    wk.sc.opt_len += 3 * ((int64_t)max_blindex + 1) + 5 + 5 + 4;
    int64_t opt_lenb = (wk.sc.opt_len + 3 + 7) >> 3, static_lenb = (wk.sc.static_len + 3 + 7) >> 3;
    if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    desc->stored_len = stored_len; desc->last = last;
    for (int i = 0; i < kHdrWords; i++) desc->hdr[i] = 0;
    HdrBits hb{desc->hdr, 0};
    if (can_store && (int64_t)desc->stored_len + 4 <= opt_lenb) {
        desc->type = 0;
        hb.put((0u << 1) + desc->last, 3);
        desc->body_bits = 0;
    } else if (static_lenb == opt_lenb) {
        desc->type = 1;
        hb.put((1u << 1) + desc->last, 3);
        desc->body_bits = (uint32_t)wk.sc.static_len;
    } else {
        desc->type = 2;
        hb.put((2u << 1) + desc->last, 3);
        hb.put((uint32_t)(wk.lt.max_code + 1 - 257), 5);
        hb.put((uint32_t)(wk.dt.max_code + 1 - 1), 5);
        hb.put((uint32_t)(max_blindex + 1 - 4), 4);
        for (int r = 0; r <= max_blindex; r++) hb.put(wk.bt.len[bl_order[r]], 3);
        tree_rle(wk.lt, wk.lt.max_code, (uint16_t *)nullptr, &wk.bt, &hb);
        tree_rle(wk.dt, wk.dt.max_code, (uint16_t *)nullptr, &wk.bt, &hb);
        desc->body_bits = (uint32_t)(wk.sc.opt_len - (hb.nbits - 3));
    }
    desc->hdr_bits = (uint32_t)hb.nbits;
}

// The code tables the bit packer reads, after build_block: (lane, nl) = (lane id, 32) on the device, (0, 1) on the host.
CCT_HD void fill_block_tables(const BlockWork &wk, BlockDesc *desc, int lane, int nl) {
    if (desc->type == 1) {
        for (int n = lane; n < kLCodes + 2; n += nl) { desc->lcode[n] = (uint16_t)static_lcode(n); desc->llen[n] = (uint8_t)static_llen(n); }
        for (int n = lane; n < kDCodes; n += nl) { desc->dcode[n] = (uint16_t)bit_reverse((uint32_t)n, 5); desc->dlen[n] = 5; }
    } else if (desc->type == 2) {
        for (int n = lane; n < kLCodes; n += nl) { desc->lcode[n] = wk.lt.code()[n]; desc->llen[n] = (uint8_t)(n <= wk.lt.max_code ? wk.lt.len[n] : 0); }
        for (int n = lane; n < kDCodes; n += nl) { desc->dcode[n] = wk.dt.code()[n]; desc->dlen[n] = (uint8_t)(n <= wk.dt.max_code ? wk.dt.len[n] : 0); }
    }
}

// bits of one LZ77 symbol under a block's code tables: returns the bit string (LSB-first, <= 48 bits)
CCT_HD uint64_t sym_bits(uint32_t sym, const uint16_t *lcode, const uint8_t *llen, const uint16_t *dcode,
                         const uint8_t *dlen, int *nbits) {
    int dist = sym_dist(sym), lc = sym_lc(sym);
    if (dist == 0) { *nbits = llen[lc]; return lcode[lc]; }
    int code = len_code(lc);
    uint64_t bits = lcode[code + 257];
    int nb = llen[code + 257];
    int extra = len_extra_bits(code);
    if (extra) { bits |= (uint64_t)(lc & ((1 << extra) - 1)) << nb; nb += extra; }
    int d = dist - 1;
    code = dist_code(d);
    bits |= (uint64_t)dcode[code] << nb; nb += dlen[code];
    extra = dist_extra_bits(code);
    if (extra) { bits |= (uint64_t)(d & ((1 << extra) - 1)) << nb; nb += extra; }
    *nbits = nb;
    return bits;
}

CCT_HD int sym_nbits(uint32_t sym, const uint8_t *llen, const uint8_t *dlen) {
    int dist = sym_dist(sym), lc = sym_lc(sym);
    if (dist == 0) return llen[lc];
    int code = len_code(lc);
    int nb = llen[code + 257] + len_extra_bits(code);
    code = dist_code(dist - 1);
    return nb + dlen[code] + dist_extra_bits(code);
}

}  // namespace cct
