// Native JSONL front-end of the streaming prover (host code, no CUDA).
//
// Reference: `stream_block_summaries_jsonl` (crates/sezkp-core/src/io_jsonl.rs:27-88) reads one serde-JSON
// `BlockSummary` (crates/sezkp-core/src/types.rs:116-151) per line, skipping blank lines, and hands the blocks to
// `ProvingBackendStream::ingest_block` one at a time (sezkp-core/src/prover.rs:104-150).  This file does the same
// for the GPU prover, but parses straight into the flat arrays of `sezkp_trace_desc` (include/sezkp_trace.h) —
// the only fields `prove_v1` reads — on several host threads: lines are independent records, so a chunk of text is
// split at line boundaries, every thread parses a contiguous run of lines into its own arrays, and the runs are
// concatenated in file order.  The grammar accepted is JSON (RFC 8259) restricted to what serde_json emits for
// this type: objects with string keys in any order, integers, `null`, arrays; unknown keys are skipped generically
// (strings with escapes, nested containers, floats), so advisory fields such as the tags cost nothing but a scan.
#include "jsonl.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <condition_variable>
#include <mutex>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace jsonl {

namespace {

// a malformed line: where it starts (the line number is only computed when the error is reported) and why
struct LineError {
    const char* line;
    std::string what;
};
struct Cursor {
    const char* p;
    const char* end;
    const char* line;  // first byte of the line being parsed
    [[noreturn]] void fail(const char* what) const { throw LineError{line, what}; }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
    }
    bool peek(char c) {
        ws();
        return p < end && *p == c;
    }
    void expect(char c) {
        ws();
        if (p >= end || *p != c) {
            char m[48];
            snprintf(m, sizeof m, "expected '%c'", c);
            fail(m);
        }
        p++;
    }
    bool accept(char c) {
        ws();
        if (p < end && *p == c) {
            p++;
            return true;
        }
        return false;
    }
    // integer in [lo, hi]; rejects fractions and exponents (serde_json rejects them for integer fields as well)
    int64_t integer(int64_t lo, uint64_t hi) {
        ws();
        bool neg = false;
        if (p < end && *p == '-') {
            neg = true;
            p++;
        }
        if (p >= end || *p < '0' || *p > '9') fail("expected an integer");
        uint64_t v = 0;
        const char* start = p;
        while (p < end && *p >= '0' && *p <= '9') {
            const uint64_t d = (uint64_t)(*p - '0');
            if (v > (UINT64_MAX - d) / 10) fail("integer out of range");
            v = v * 10 + d;
            p++;
        }
        if (p - start > 1 && *start == '0') fail("leading zero in a number");
        if (p < end && (*p == '.' || *p == 'e' || *p == 'E')) fail("expected an integer, found a float");
        if (neg) {
            if (v > (uint64_t)INT64_MAX + 1) fail("integer out of range");
            const int64_t val = v == 0 ? 0 : -(int64_t)(v - 1) - 1;
            if (val < lo) fail("integer out of range");
            return val;
        }
        if (v > hi) fail("integer out of range");
        return (int64_t)v;
    }
    // string token -> [s, e) of its raw contents (keys of this schema never contain escapes; an escaped key just
    // fails to match and is skipped as unknown)
    void string_raw(const char*& s, const char*& e) {
        expect('"');
        s = p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) fail("unterminated string");
            }
            p++;
        }
        if (p >= end) fail("unterminated string");
        e = p;
        p++;
    }
    bool null_lit() {
        ws();
        if (end - p >= 4 && std::memcmp(p, "null", 4) == 0) {
            p += 4;
            return true;
        }
        return false;
    }
    void skip_value(int depth = 0) {
        if (depth > 64) fail("nesting too deep");
        ws();
        if (p >= end) fail("unexpected end of line");
        const char c = *p;
        if (c == '"') {
            const char *s, *e;
            string_raw(s, e);
        } else if (c == '{') {
            p++;
            if (accept('}')) return;
            do {
                const char *s, *e;
                string_raw(s, e);
                expect(':');
                skip_value(depth + 1);
            } while (accept(','));
            expect('}');
        } else if (c == '[') {
            p++;
            if (accept(']')) return;
            do skip_value(depth + 1);
            while (accept(','));
            expect(']');
        } else if (c == 't' && end - p >= 4 && std::memcmp(p, "true", 4) == 0) {
            p += 4;
        } else if (c == 'f' && end - p >= 5 && std::memcmp(p, "false", 5) == 0) {
            p += 5;
        } else if (c == 'n' && end - p >= 4 && std::memcmp(p, "null", 4) == 0) {
            p += 4;
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            p++;
            while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) p++;
        } else {
            fail("unexpected character");
        }
    }
};

inline bool key_is(const char* s, const char* e, const char* name) {
    const size_t n = std::strlen(name);
    return (size_t)(e - s) == n && std::memcmp(s, name, n) == 0;
}

// {"write": null | u16, "mv": i8}
// Fast path for the exact bytes serde_json emits (`{"write":null,"mv":-1}`, no whitespace, struct field order); any
// deviation falls through to the general parser below, which also produces the error messages.
inline bool fast_small_int(const char*& q, const char* end, int64_t lo, int64_t hi, int64_t& out) {
    const char* p = q;
    bool neg = false;
    if (p < end && *p == '-') {
        neg = true;
        p++;
    }
    if (p >= end || *p < '0' || *p > '9') return false;
    if (*p == '0' && p + 1 < end && p[1] >= '0' && p[1] <= '9') return false;  // leading zero: let the general path reject it
    int64_t v = 0;
    int digits = 0;
    while (p < end && *p >= '0' && *p <= '9' && digits < 6) {
        v = v * 10 + (*p - '0');
        p++;
        digits++;
    }
    if (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E')) return false;
    if (neg) v = -v;
    if (v < lo || v > hi) return false;
    out = v;
    q = p;
    return true;
}
inline bool fast_tape_op(Cursor& c, int8_t& mv, uint8_t& wf, uint16_t& ws) {
    const char* q = c.p;
    const char* const end = c.end;
    if (end - q < 20 || std::memcmp(q, "{\"write\":", 9) != 0) return false;
    q += 9;
    int64_t v;
    if (*q == 'n') {
        if (std::memcmp(q, "null", 4) != 0) return false;
        q += 4;
        wf = 0;
        ws = 0;
    } else {
        if (!fast_small_int(q, end, 0, 65535, v)) return false;
        wf = 1;
        ws = (uint16_t)v;
    }
    if (end - q < 8 || std::memcmp(q, ",\"mv\":", 6) != 0) return false;
    q += 6;
    if (!fast_small_int(q, end, -128, 127, v)) return false;
    if (q >= end || *q != '}') return false;
    mv = (int8_t)v;
    c.p = q + 1;
    return true;
}

void parse_tape_op(Cursor& c, int8_t& mv, uint8_t& wf, uint16_t& ws) {
    if (fast_tape_op(c, mv, wf, ws)) return;
    bool have_mv = false, have_w = false;
    c.expect('{');
    if (!c.accept('}')) {
        do {
            const char *s, *e;
            c.string_raw(s, e);
            c.expect(':');
            if (key_is(s, e, "mv")) {
                mv = (int8_t)c.integer(-128, 127);
                have_mv = true;
            } else if (key_is(s, e, "write")) {
                if (c.null_lit()) {
                    wf = 0;
                    ws = 0;
                } else {
                    wf = 1;
                    ws = (uint16_t)c.integer(0, 65535);
                }
                have_w = true;
            } else {
                c.skip_value();
            }
        } while (c.accept(','));
        c.expect('}');
    }
    if (!have_mv || !have_w) c.fail("tape op needs \"write\" and \"mv\"");
}

// {"input_mv": i8, "tapes": [TapeOp; tau]}
// Fast path for `{"input_mv":-1,"tapes":[{...},...]}` exactly as serde_json writes it; on any deviation the vectors are
// rolled back and the general parser below takes the step again.
inline bool fast_step(Cursor& c, Trace& t, uint32_t& tau) {
    const char* q = c.p;
    const char* const end = c.end;
    if (end - q < 24 || std::memcmp(q, "{\"input_mv\":", 12) != 0) return false;
    q += 12;
    int64_t v;
    if (!fast_small_int(q, end, -128, 127, v)) return false;
    if (end - q < 11 || std::memcmp(q, ",\"tapes\":[", 10) != 0) return false;
    q += 10;
    const size_t n0 = t.mv.size();
    uint32_t r = 0;
    Cursor cc{q, end, c.line};
    bool ok;
    if (tau != 0) {  // known width: grow the three arrays once per step and write through raw pointers
        t.mv.resize(n0 + tau);
        t.write_flag.resize(n0 + tau);
        t.write_sym.resize(n0 + tau);
        int8_t* pm = t.mv.data() + n0;
        uint8_t* pf = t.write_flag.data() + n0;
        uint16_t* ps = t.write_sym.data() + n0;
        ok = true;
        for (; r < tau; r++) {
            if (r && !(cc.p < end && *cc.p++ == ',')) {
                ok = false;
                break;
            }
            if (!fast_tape_op(cc, pm[r], pf[r], ps[r])) {
                ok = false;
                break;
            }
        }
        ok = ok && end - cc.p >= 2 && cc.p[0] == ']' && cc.p[1] == '}';
    } else {
        for (;;) {
            int8_t mv = 0;
            uint8_t wf = 0;
            uint16_t ws = 0;
            if (!fast_tape_op(cc, mv, wf, ws)) break;
            t.mv.push_back(mv);
            t.write_flag.push_back(wf);
            t.write_sym.push_back(ws);
            r++;
            if (cc.p < end && *cc.p == ',') {
                cc.p++;
                continue;
            }
            break;
        }
        ok = r > 0 && end - cc.p >= 2 && cc.p[0] == ']' && cc.p[1] == '}';
    }
    if (!ok) {
        t.mv.resize(n0);
        t.write_flag.resize(n0);
        t.write_sym.resize(n0);
        return false;
    }
    if (tau == 0) tau = r;
    t.input_mv.push_back((int8_t)v);
    c.p = cc.p + 2;
    return true;
}

// Bulk fast path for the steps array (the bytes that make up 97 % of a line): as many consecutive steps as are written
// exactly the way serde_json writes them — `{"input_mv":-1,"tapes":[{"write":null,"mv":1},...tau ops...]}` separated by
// single commas — are parsed with fixed-width loads and compares, no per-byte bounds checks (every tape op starts with
// at least SLACK readable bytes ahead, longer than the longest op this path accepts) and raw stores into arrays grown
// once per call.  Returns the number of steps taken; the cursor is left right behind the last one (at ',' or ']').
// Anything else — whitespace, other key order, large numbers, the last bytes of a line — is left to parse_step, which
// also produces the error messages.
inline uint64_t ld8(const char* p) {
    uint64_t v;
    std::memcpy(&v, p, 8);
    return v;
}
inline uint32_t ld4(const char* p) {
    uint32_t v;
    std::memcpy(&v, p, 4);
    return v;
}
constexpr uint64_t lit8(const char (&s)[9]) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; i--) v = (v << 8) | (uint8_t)s[i];
    return v;
}
constexpr uint32_t lit4(const char (&s)[5]) {
    uint32_t v = 0;
    for (int i = 3; i >= 0; i--) v = (v << 8) | (uint8_t)s[i];
    return v;
}
// unsigned decimal of 1..5 digits without a leading zero, not followed by a digit, '.', 'e' or 'E'
inline bool fast_uint(const char*& q, uint32_t& out) {
    uint32_t d = (uint32_t)(uint8_t)q[0] - '0';
    if (d > 9) return false;
    const char* p = q + 1;
    uint32_t v = d, e = (uint32_t)(uint8_t)*p - '0';
    if (e <= 9) {
        if (d == 0) return false;
        int digits = 1;
        do {
            v = v * 10 + e;
            p++;
            e = (uint32_t)(uint8_t)*p - '0';
        } while (e <= 9 && ++digits < 5);
        if (e <= 9) return false;
    }
    if (*p == '.' || *p == 'e' || *p == 'E') return false;
    out = v;
    q = p;
    return true;
}
// One tape op behind its `{"write":` in the less common shapes (3-5 digit symbols, multi-digit moves); s has FAST_SLACK
// readable bytes behind the op's first byte.  Out of line so that the common path of fast_tapes stays small.
__attribute__((noinline)) const char* slow_tape(const char* s, int8_t* pm, uint8_t* pf, uint16_t* ps) {
    constexpr uint64_t MV6 = lit8(",\"mv\":\0\0"), M6 = 0xFFFFFFFFFFFFull;
    uint32_t v;
    if (ld4(s) == lit4("null")) {
        s += 4;
        *pf = 0;
        *ps = 0;
    } else {
        if (!fast_uint(s, v) || v > 65535) return nullptr;
        *pf = 1;
        *ps = (uint16_t)v;
    }
    if ((ld8(s) & M6) != MV6) return nullptr;
    s += 6;
    const bool neg = *s == '-';
    s += neg;
    if (!fast_uint(s, v) || v > 127u + neg) return nullptr;
    if (*s != '}') return nullptr;
    *pm = (int8_t)(neg ? -(int32_t)v : (int32_t)v);
    return s + 1;
}
// The tau tape ops of one step (see fast_steps); returns the position behind the last op, or null when the bytes are not
// in the exact serde form.  Kept out of line: the decode wants its dozen temporaries in registers, and inlined into
// fast_steps the compiler spilled them.
constexpr ptrdiff_t FAST_SLACK = 40;  // > 9 + 5 + 6 + 4 + 1 bytes of the longest accepted tape op + the separators after it
__attribute__((noinline)) const char* fast_tapes(const char* s, const char* const end, const uint32_t tau, int8_t* pm, uint8_t* pf,
                                                 uint16_t* ps) {
    for (uint32_t r = 0; r < tau; r++) {
        if (end - s < FAST_SLACK) return nullptr;
        if (r) {
            if (*s != ',') return nullptr;
            s++;
        }
        // {"write":
        if (ld8(s) != lit8("{\"write\"") || s[8] != ':') return nullptr;
        s += 9;
        constexpr uint64_t MV6 = lit8(",\"mv\":\0\0"), M6 = 0xFFFFFFFFFFFFull;
#if defined(__SSE2__)
        {   // common shapes — write = null | D | DD, mv = D | -D — without data-dependent branches (null or number, one
            // or two digits and the sign are coin flips in real traces): the closing brace is found with one 16-byte
            // compare, every field position follows from it, and only the next op's address depends on the search.
            const unsigned mask = (unsigned)_mm_movemask_epi8(_mm_cmpeq_epi8(_mm_loadu_si128((const __m128i*)s), _mm_set1_epi8('}')));
            const unsigned pos = (unsigned)__builtin_ctz(mask | 0x10000u);  // s[pos] == '}' (16: none in range)
            if (pos >= 8 && pos < 16) {
                const uint32_t dm = (uint32_t)(uint8_t)s[pos - 1] - '0';
                const uint32_t ng = s[pos - 2] == '-';
                const uint32_t wlen = pos - 7 - ng;  // bytes of the write field
                const uint32_t d0 = (uint32_t)(uint8_t)s[0] - '0', d1 = (uint32_t)(uint8_t)s[1] - '0';
                const uint32_t isnull = (wlen == 4) & (ld4(s) == lit4("null"));
                const uint32_t one = (wlen == 1) & (d0 <= 9), two = (wlen == 2) & (d0 - 1 <= 8) & (d1 <= 9);
                if (((ld8(s + wlen) & M6) == MV6) & (dm <= 9) & (isnull | one | two)) {
                    pf[r] = (uint8_t)(isnull ^ 1);
                    ps[r] = (uint16_t)((d0 + ((0u - two) & (9 * d0 + d1))) & (0u - (isnull ^ 1)));
                    pm[r] = (int8_t)((int32_t)(dm ^ (0u - ng)) + (int32_t)ng);
                    s += pos + 1;
                    continue;
                }
            }
        }
#endif
        s = slow_tape(s, pm + r, pf + r, ps + r);
        if (!s) return nullptr;
    }
    return s;
}
size_t fast_steps(Cursor& c, Trace& t, uint32_t tau) {
    constexpr ptrdiff_t SLACK = FAST_SLACK;
    const char* q = c.p;
    const char* const end = c.end;
    const size_t min_step = 25 + (size_t)tau * 19;
    if ((size_t)(end - q) < min_step) return 0;
    const size_t cap_steps = (size_t)(end - q) / min_step + 1;
    const size_t rows0 = t.input_mv.size(), cells0 = t.mv.size();
    t.input_mv.resize(rows0 + cap_steps);
    t.mv.resize(cells0 + cap_steps * tau);
    t.write_flag.resize(cells0 + cap_steps * tau);
    t.write_sym.resize(cells0 + cap_steps * tau);
    int8_t* pi = t.input_mv.data() + rows0;
    int8_t* pm = t.mv.data() + cells0;
    uint8_t* pf = t.write_flag.data() + cells0;
    uint16_t* ps = t.write_sym.data() + cells0;
    size_t done = 0;
    for (;;) {
        const char* s = q;
        if (done) {
            if (*s != ',') break;
            s++;
        }
        if (end - s < SLACK) break;
        // {"input_mv":
        if (ld8(s) != lit8("{\"input_") || ld4(s + 8) != lit4("mv\":")) break;
        s += 12;
        bool neg = *s == '-';
        s += neg;
        uint32_t v;
        if (!fast_uint(s, v) || v > 127u + neg) break;
        const int8_t in_mv = (int8_t)(neg ? -(int32_t)v : (int32_t)v);
        // ,"tapes":[
        if (ld8(s) != lit8(",\"tapes\"") || s[8] != ':' || s[9] != '[') break;
        s += 10;
        s = fast_tapes(s, end, tau, pm, pf, ps);
        if (!s || end - s < 2 || s[0] != ']' || s[1] != '}') break;
        q = s + 2;
        pi[done] = in_mv;
        pm += tau;
        pf += tau;
        ps += tau;
        done++;
        if (done == cap_steps) break;
    }
    t.input_mv.resize(rows0 + done);
    t.mv.resize(cells0 + done * tau);
    t.write_flag.resize(cells0 + done * tau);
    t.write_sym.resize(cells0 + done * tau);
    c.p = q;
    return done;
}

void parse_step(Cursor& c, Trace& t, uint32_t& tau) {
    if (fast_step(c, t, tau)) return;
    bool have_in = false, have_tapes = false;
    int8_t in_mv = 0;
    c.expect('{');
    if (!c.accept('}')) {
        do {
            const char *s, *e;
            c.string_raw(s, e);
            c.expect(':');
            if (key_is(s, e, "input_mv")) {
                in_mv = (int8_t)c.integer(-128, 127);
                have_in = true;
            } else if (key_is(s, e, "tapes")) {
                if (have_tapes) c.fail("duplicate \"tapes\"");
                uint32_t r = 0;
                c.expect('[');
                if (!c.accept(']')) {
                    do {
                        int8_t mv = 0;
                        uint8_t wf = 0;
                        uint16_t ws = 0;
                        parse_tape_op(c, mv, wf, ws);
                        t.mv.push_back(mv);
                        t.write_flag.push_back(wf);
                        t.write_sym.push_back(ws);
                        r++;
                    } while (c.accept(','));
                    c.expect(']');
                }
                if (tau == 0) tau = r;
                if (r != tau || r == 0) c.fail("step has a different number of tapes than the first block's windows");
                have_tapes = true;
            } else {
                c.skip_value();
            }
        } while (c.accept(','));
        c.expect('}');
    }
    if (!have_in || !have_tapes) c.fail("step needs \"input_mv\" and \"tapes\"");
    t.input_mv.push_back(in_mv);
}

template <class T, class F>
uint32_t parse_array(Cursor& c, std::vector<T>& out, F&& elem) {
    uint32_t n = 0;
    c.expect('[');
    if (c.accept(']')) return 0;
    do {
        out.push_back(elem());
        n++;
    } while (c.accept(','));
    c.expect(']');
    return n;
}

// one BlockSummary object; appends to t.  tau: 0 = not known yet (taken from this block's windows)
void parse_block(Cursor& c, Trace& t, uint32_t& tau) {
    enum { STEP_LO = 1, STEP_HI = 2, WINDOWS = 4, IN_OFF = 8, OUT_OFF = 16, LOG = 32 };
    unsigned seen = 0;
    uint64_t step_lo = 0, step_hi = 0, n_steps = 0;
    uint32_t n_win = 0, n_in = 0, n_out = 0;
    Manifest mf{};
    const size_t rows_before = t.input_mv.size();
    c.expect('{');
    if (!c.accept('}')) {
        do {
            const char *s, *e;
            c.string_raw(s, e);
            c.expect(':');
            if (key_is(s, e, "step_lo")) {
                step_lo = (uint64_t)c.integer(0, UINT64_MAX >> 1);
                seen |= STEP_LO;
            } else if (key_is(s, e, "step_hi")) {
                step_hi = (uint64_t)c.integer(0, UINT64_MAX >> 1);
                seen |= STEP_HI;
            } else if (key_is(s, e, "version")) {
                mf.version = (uint16_t)c.integer(0, 65535);
            } else if (key_is(s, e, "block_id")) {
                mf.block_id = (uint32_t)c.integer(0, UINT32_MAX);
            } else if (key_is(s, e, "ctrl_in")) {
                mf.ctrl_in = (uint16_t)c.integer(0, 65535);
            } else if (key_is(s, e, "ctrl_out")) {
                mf.ctrl_out = (uint16_t)c.integer(0, 65535);
            } else if (key_is(s, e, "in_head_in")) {
                mf.in_head_in = c.integer(INT64_MIN, INT64_MAX);
            } else if (key_is(s, e, "in_head_out")) {
                mf.in_head_out = c.integer(INT64_MIN, INT64_MAX);
            } else if (key_is(s, e, "windows")) {
                if (seen & WINDOWS) c.fail("duplicate \"windows\"");
                c.expect('[');
                if (!c.accept(']')) {
                    do {
                        bool hl = false, hr = false;
                        int64_t l = 0, r = 0;
                        c.expect('{');
                        if (!c.accept('}')) {
                            do {
                                const char *ks, *ke;
                                c.string_raw(ks, ke);
                                c.expect(':');
                                if (key_is(ks, ke, "left")) {
                                    l = c.integer(INT64_MIN, INT64_MAX);
                                    hl = true;
                                } else if (key_is(ks, ke, "right")) {
                                    r = c.integer(INT64_MIN, INT64_MAX);
                                    hr = true;
                                } else {
                                    c.skip_value();
                                }
                            } while (c.accept(','));
                            c.expect('}');
                        }
                        if (!hl || !hr) c.fail("window needs \"left\" and \"right\"");
                        t.win_left.push_back(l);
                        t.win_right.push_back(r);
                        n_win++;
                    } while (c.accept(','));
                    c.expect(']');
                }
                seen |= WINDOWS;
            } else if (key_is(s, e, "head_in_offsets")) {
                if (seen & IN_OFF) c.fail("duplicate \"head_in_offsets\"");
                n_in = parse_array(c, t.head_in_off, [&] { return (uint32_t)c.integer(0, UINT32_MAX); });
                seen |= IN_OFF;
            } else if (key_is(s, e, "head_out_offsets")) {
                if (seen & OUT_OFF) c.fail("duplicate \"head_out_offsets\"");
                n_out = parse_array(c, t.head_out_off, [&] { return (uint32_t)c.integer(0, UINT32_MAX); });
                seen |= OUT_OFF;
            } else if (key_is(s, e, "movement_log")) {
                if (seen & LOG) c.fail("duplicate \"movement_log\"");
                bool have_steps = false;
                // the steps may precede "windows" in the object: tau is then fixed by the first step and checked below
                c.expect('{');
                if (!c.accept('}')) {
                    do {
                        const char *ks, *ke;
                        c.string_raw(ks, ke);
                        c.expect(':');
                        if (key_is(ks, ke, "steps")) {
                            c.expect('[');
                            if (!c.accept(']')) {
                                do {
                                    const size_t k = tau ? fast_steps(c, t, tau) : 0;
                                    n_steps += k;
                                    if (k == 0) {
                                        parse_step(c, t, tau);
                                        n_steps++;
                                    }
                                } while (c.accept(','));
                                c.expect(']');
                            }
                            have_steps = true;
                        } else {
                            c.skip_value();
                        }
                    } while (c.accept(','));
                    c.expect('}');
                }
                if (!have_steps) c.fail("movement_log needs \"steps\"");
                seen |= LOG;
            } else {
                c.skip_value();  // pre_tags, post_tags, anything a newer schema adds
            }
        } while (c.accept(','));
        c.expect('}');
    }
    c.ws();
    if (c.p != c.end) c.fail("trailing characters after the block object");
    if (seen != (STEP_LO | STEP_HI | WINDOWS | IN_OFF | OUT_OFF | LOG)) c.fail("block is missing a required field");
    if (tau == 0) tau = n_win;
    if (n_win != tau || n_in != tau || n_out != tau || tau == 0) c.fail("windows / head offsets do not have tau entries");
    if (step_hi < step_lo || step_hi - step_lo + 1 != n_steps) c.fail("movement_log length differs from step_hi - step_lo + 1");
    if (t.input_mv.size() - rows_before != n_steps) c.fail("internal: step count mismatch");
    mf.step_lo = step_lo;
    mf.step_hi = step_hi;
    t.manifest.push_back(mf);
    t.block_len.push_back(n_steps);
}

void append(Trace& dst, const Trace& src) {
    auto cat = [](auto& a, const auto& b) { a.insert(a.end(), b.begin(), b.end()); };
    cat(dst.block_len, src.block_len);
    cat(dst.win_left, src.win_left);
    cat(dst.win_right, src.win_right);
    cat(dst.head_in_off, src.head_in_off);
    cat(dst.head_out_off, src.head_out_off);
    cat(dst.input_mv, src.input_mv);
    cat(dst.mv, src.mv);
    cat(dst.write_flag, src.write_flag);
    cat(dst.write_sym, src.write_sym);
    cat(dst.manifest, src.manifest);
}

}  // namespace

void Trace::clear_keep_capacity() {
    tau = 0;
    n_lines = 0;
    block_len.clear();
    win_left.clear();
    win_right.clear();
    head_in_off.clear();
    head_out_off.clear();
    input_mv.clear();
    mv.clear();
    write_flag.clear();
    write_sym.clear();
    manifest.clear();
}
void Trace::fill_desc(sezkp_trace_desc& d) const {
    d.tau = tau;
    d.flags = 0;
    d.n_blocks = block_len.size();
    d.n_rows = input_mv.size();
    d.block_len = block_len.data();
    d.win_left = win_left.data();
    d.win_right = win_right.data();
    d.head_in_off = head_in_off.data();
    d.head_out_off = head_out_off.data();
    d.input_mv = input_mv.data();
    d.mv = mv.data();
    d.write_flag = write_flag.data();
    d.write_sym = write_sym.data();
}

// Parse every non-blank line of text[0, len) (the last line may lack its '\n') into `parts`, one Trace per worker
// in file order.  No pass over the text is serial: worker i owns the lines whose first byte lies in
// [len*i/T, len*(i+1)/T), finds its first line start itself, and counts the newlines of its range on the way.
// tau_hint = 0: every worker takes tau from its first block and the parts are checked against each other.
// first_line_no is only used in error messages.  Returns the number of lines seen (blank ones included).
struct WorkerPool::Impl {
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> th;
    const std::function<void(int)>* fn = nullptr;
    int tasks = 0, next = 0, pending = 0;
    uint64_t gen = 0;
    bool quit = false;
    // take tasks until none is left; called with the lock held, returns with it held
    void drain(std::unique_lock<std::mutex>& lk) {
        while (next < tasks) {
            const int i = next++;
            const std::function<void(int)>* f = fn;
            lk.unlock();
            (*f)(i);
            lk.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    void worker() {
        std::unique_lock<std::mutex> lk(mu);
        uint64_t seen = 0;
        for (;;) {
            cv_job.wait(lk, [&] { return quit || gen != seen; });
            if (quit) return;
            seen = gen;
            drain(lk);
        }
    }
};
WorkerPool::WorkerPool(int n_threads) : p_(new Impl()), n_(std::max(1, std::min(n_threads, 256))) {
    try {
        for (int i = 1; i < n_; i++) p_->th.emplace_back([this] { p_->worker(); });  // the caller of run() is the n-th worker
    } catch (...) {
        n_ = (int)p_->th.size() + 1;  // fewer threads than asked for: still correct
    }
}
WorkerPool::~WorkerPool() {
    {
        std::lock_guard<std::mutex> lk(p_->mu);
        p_->quit = true;
    }
    p_->cv_job.notify_all();
    for (auto& t : p_->th) t.join();
    delete p_;
}
void WorkerPool::run(int tasks, const std::function<void(int)>& fn) {
    if (tasks <= 0) return;
    std::unique_lock<std::mutex> lk(p_->mu);
    p_->fn = &fn;
    p_->tasks = tasks;
    p_->next = 0;
    p_->pending = tasks;
    p_->gen++;
    p_->cv_job.notify_all();
    p_->drain(lk);
    p_->cv_done.wait(lk, [&] { return p_->pending == 0; });
    p_->fn = nullptr;
    p_->tasks = p_->next = 0;
}

size_t parse_parts(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, std::vector<Trace>& parts,
                   uint32_t& tau_out, WorkerPool* pool) {
    const char* const end = text + len;
    int T = std::max(1, std::min(n_threads, 256));
    if (len < ((size_t)T << 16)) T = (int)std::max<size_t>(1, len >> 16);  // at least 64 KiB of text per worker
    parts.resize((size_t)T);  // elements a caller left in place are reused (their arrays keep their capacity)
    for (auto& t : parts) t.clear_keep_capacity();
    struct Result {
        const char* err_line = nullptr;
        std::string err;
        size_t newlines = 0;
        bool oom = false;
    };
    std::vector<Result> res((size_t)T);
    auto work = [&](int ti) {
        const char* lo = text + len * (size_t)ti / (size_t)T;
        const char* const hi = text + len * (size_t)(ti + 1) / (size_t)T;
        Trace& t = parts[(size_t)ti];
        Result& r = res[(size_t)ti];
        t.tau = tau_hint;
        // a line belongs to the worker in whose range its first byte lies
        if (ti > 0 && lo[-1] != '\n') {
            const char* nl = (const char*)std::memchr(lo, '\n', (size_t)(end - lo));
            lo = nl ? nl + 1 : end;
        }
        try {
            const size_t cells = (size_t)(hi > lo ? hi - lo : 0) / 20 + 16;  // ~26 bytes of JSON per (row, tape) cell
            t.mv.reserve(cells);
            t.write_flag.reserve(cells);
            t.write_sym.reserve(cells);
            t.input_mv.reserve(cells / 2 + 16);
            uint32_t tau = tau_hint;
            const char* p = lo;
            while (p < hi) {
                const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
                const char* e = nl ? nl : end;
                const char *b2 = p, *e2 = e;
                while (b2 < e2 && (*b2 == ' ' || *b2 == '\t' || *b2 == '\r')) b2++;
                while (e2 > b2 && (e2[-1] == ' ' || e2[-1] == '\t' || e2[-1] == '\r')) e2--;
                if (e2 > b2) {  // blank lines are skipped (io_jsonl.rs:52-55)
                    Cursor c{b2, e2, p};
                    parse_block(c, t, tau);
                }
                r.newlines++;
                p = nl ? nl + 1 : end;
            }
            t.tau = tau;
        } catch (const LineError& le) {
            r.err_line = le.line;
            r.err = le.what;
        } catch (const std::bad_alloc&) {
            r.oom = true;
        }
    };
    if (T == 1) {
        work(0);
    } else if (pool) {
        pool->run(T, work);
    } else {
        std::vector<std::thread> th;
        for (int ti = 0; ti < T; ti++) th.emplace_back(work, ti);
        for (auto& x : th) x.join();
    }
    auto report = [&](const char* line, const std::string& what) {
        size_t no = first_line_no;
        for (const char* q = text; q < line;) {
            const char* x = (const char*)std::memchr(q, '\n', (size_t)(line - q));
            if (!x) break;
            no++;
            q = x + 1;
        }
        char buf[200];
        snprintf(buf, sizeof buf, "jsonl line %zu: %s", no, what.c_str());
        throw std::runtime_error(buf);
    };
    size_t lines = 0;
    uint32_t tau = tau_hint;
    for (int ti = 0; ti < T; ti++) {  // the first failure in file order wins
        if (res[(size_t)ti].oom) throw std::bad_alloc();
        if (res[(size_t)ti].err_line) report(res[(size_t)ti].err_line, res[(size_t)ti].err);
        lines += res[(size_t)ti].newlines;
        const Trace& t = parts[(size_t)ti];
        if (!t.block_len.empty()) {
            if (tau == 0) tau = t.tau;
            if (t.tau != tau) {  // blame the first line of the part that disagrees
                const char* lo = text + len * (size_t)ti / (size_t)T;
                if (ti > 0 && lo[-1] != '\n') {
                    const char* nl = (const char*)std::memchr(lo, '\n', (size_t)(end - lo));
                    lo = nl ? nl + 1 : end;
                }
                report(lo, "block has a different number of tapes (tau) than the blocks before it");
            }
        }
    }
    tau_out = tau;
    return lines;
}

void parse(const char* text, size_t len, int n_threads, uint32_t tau_hint, size_t first_line_no, Trace& out) {
    std::vector<Trace> parts;
    uint32_t tau = tau_hint;
    const size_t lines = parse_parts(text, len, n_threads, tau_hint, first_line_no, parts, tau);
    if (parts.size() == 1) {
        out = std::move(parts[0]);
    } else {
        out = Trace();
        size_t rows = 0, blocks = 0;
        for (auto& p : parts) {
            rows += p.input_mv.size();
            blocks += p.block_len.size();
        }
        out.block_len.reserve(blocks);
        out.manifest.reserve(blocks);
        out.win_left.reserve(blocks * tau);
        out.win_right.reserve(blocks * tau);
        out.head_in_off.reserve(blocks * tau);
        out.head_out_off.reserve(blocks * tau);
        out.input_mv.reserve(rows);
        out.mv.reserve(rows * tau);
        out.write_flag.reserve(rows * tau);
        out.write_sym.reserve(rows * tau);
        for (auto& p : parts) {
            append(out, p);
            p = Trace();
        }
    }
    out.tau = tau;
    out.n_lines = lines;
}

// ---- writer: the inverse of the parser, the text serde_json::to_string emits for a BlockSummary (field order of
//      crates/sezkp-core/src/types.rs:116-151; what the reference CLI's `export-jsonl` writes, one block per line) ----
namespace {
inline void put_str(std::string& o, const char* s) { o.append(s); }
inline void put_int(std::string& o, long long v) {
    char buf[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do {
        buf[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) o.push_back('-');
    while (n) o.push_back(buf[--n]);
}
void format_block(std::string& o, const sezkp_trace_desc& d, const sezkp_block_scalars* sc, uint64_t k, uint64_t row0, const std::string& tags) {
    const uint64_t n = d.block_len[k], tau = d.tau;
    put_str(o, "{\"version\":");
    put_int(o, sc ? sc[k].version : 1);
    put_str(o, ",\"block_id\":");
    put_int(o, sc ? sc[k].block_id : (long long)(k + 1));
    put_str(o, ",\"step_lo\":");
    put_int(o, sc ? (long long)sc[k].step_lo : (long long)(row0 + 1));
    put_str(o, ",\"step_hi\":");
    put_int(o, sc ? (long long)sc[k].step_hi : (long long)(row0 + n));
    put_str(o, ",\"ctrl_in\":");
    put_int(o, sc ? sc[k].ctrl_in : 0);
    put_str(o, ",\"ctrl_out\":");
    put_int(o, sc ? sc[k].ctrl_out : 0);
    put_str(o, ",\"in_head_in\":");
    put_int(o, sc ? sc[k].in_head_in : 0);
    put_str(o, ",\"in_head_out\":");
    put_int(o, sc ? sc[k].in_head_out : 0);
    put_str(o, ",\"windows\":[");
    for (uint64_t r = 0; r < tau; r++) {
        put_str(o, r ? ",{\"left\":" : "{\"left\":");
        put_int(o, d.win_left[k * tau + r]);
        put_str(o, ",\"right\":");
        put_int(o, d.win_right[k * tau + r]);
        o.push_back('}');
    }
    put_str(o, "],\"head_in_offsets\":[");
    for (uint64_t r = 0; r < tau; r++) {
        if (r) o.push_back(',');
        put_int(o, d.head_in_off[k * tau + r]);
    }
    put_str(o, "],\"head_out_offsets\":[");
    for (uint64_t r = 0; r < tau; r++) {
        if (r) o.push_back(',');
        put_int(o, d.head_out_off[k * tau + r]);
    }
    put_str(o, "],\"movement_log\":{\"steps\":[");
    for (uint64_t j = 0; j < n; j++) {
        const uint64_t i = row0 + j;
        put_str(o, j ? ",{\"input_mv\":" : "{\"input_mv\":");
        put_int(o, d.input_mv[i]);
        put_str(o, ",\"tapes\":[");
        for (uint64_t r = 0; r < tau; r++) {
            put_str(o, r ? ",{\"write\":" : "{\"write\":");
            if (d.write_flag[i * tau + r]) put_int(o, d.write_sym[i * tau + r]);
            else put_str(o, "null");
            put_str(o, ",\"mv\":");
            put_int(o, d.mv[i * tau + r]);
            o.push_back('}');
        }
        put_str(o, "]}");
    }
    put_str(o, "]},\"pre_tags\":");
    o.append(tags);
    put_str(o, ",\"post_tags\":");
    o.append(tags);
    put_str(o, "}\n");
}
}  // namespace

size_t write_file(const char* path, const sezkp_trace_desc& d, const sezkp_block_scalars* scalars, int n_threads) {
    if (d.flags != 0) throw std::runtime_error("jsonl writer: packed descriptors are not supported");
    FILE* f = std::fopen(path, "wb");
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::string tags = "[";
    for (uint32_t r = 0; r < d.tau; r++) tags += r ? ",[0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0]" : "[0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0]";
    tags += "]";
    std::vector<uint64_t> start(d.n_blocks + 1, 0);
    for (uint64_t k = 0; k < d.n_blocks; k++) start[k + 1] = start[k] + d.block_len[k];
    if (n_threads < 1) n_threads = 1;
    const uint64_t batch = 256 * (uint64_t)n_threads;  // blocks formatted per round (bounded memory: ~100 KB of text per block)
    std::vector<std::string> bufs(n_threads);
    size_t total = 0;
    bool ok = true;
    for (uint64_t k0 = 0; k0 < d.n_blocks && ok; k0 += batch) {
        const uint64_t k1 = std::min<uint64_t>(d.n_blocks, k0 + batch), per = (k1 - k0 + n_threads - 1) / n_threads;
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++)
            th.emplace_back([&, t] {
                std::string& o = bufs[t];
                o.clear();
                for (uint64_t k = k0 + t * per; k < std::min(k1, k0 + (t + 1) * per); k++) format_block(o, d, scalars, k, start[k], tags);
            });
        for (auto& x : th) x.join();
        for (int t = 0; t < n_threads && ok; t++) {
            ok = std::fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
            total += bufs[t].size();
        }
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error(std::string("write error on ") + path);
    return total;
}

}  // namespace jsonl
