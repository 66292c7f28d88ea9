// Sanitizer harness of the native JSONL parser (csrc/jsonl.cpp): every truncation point near both ends of a small file and
// a sample in between, plus single-byte corruptions, parsed from exact-size heap buffers under ASan + UBSan — the bulk fast path
// reads with fixed-width loads, so a read past the end of the text would show up here.  Run by tests/test_jsonl_cpu.py.
#include "jsonl.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
int main(int argc, char** argv) {
    std::ifstream f(argv[1], std::ios::binary);
    std::stringstream ss; ss << f.rdbuf();
    std::string s = ss.str();
    // keep the first 8 lines; exact-size heap buffers so that ASan sees any read past the end
    size_t pos = 0; for (int i = 0; i < 8; i++) pos = s.find('\n', pos) + 1;
    s.resize(pos);
    size_t ok = 0, bad = 0;
    const bool pool_only = argc > 2 && std::string(argv[2]) == "pool";
    if (pool_only) s.clear();
    // every truncation point in the last 300 bytes and a sample elsewhere, with and without the trailing newline
    for (size_t cut = 1; cut <= s.size(); cut += (cut + 400 > s.size() || cut < 400) ? 1 : 9973) {
        char* buf = (char*)malloc(cut);
        memcpy(buf, s.data(), cut);
        for (int th = 1; th <= 3; th += 2) {
            try { jsonl::Trace t; jsonl::parse(buf, cut, th, 0, 1, t); ok++; } catch (const std::runtime_error&) { bad++; }
        }
        free(buf);
    }
    // mutations: flip one byte at a stride
    for (size_t i = 0; i < s.size(); i += 131) {
        char* buf = (char*)malloc(s.size());
        memcpy(buf, s.data(), s.size());
        buf[i] ^= 0x15;
        try { jsonl::Trace t; jsonl::parse(buf, s.size(), 1, 0, 1, t); ok++; } catch (const std::runtime_error&) { bad++; }
        free(buf);
    }
    printf("parsed ok %zu, rejected %zu\n", ok, bad);
    // The file path of the prover (capi.cu: sezkp_stark_v1_prove_jsonl_file) parses a file piece by piece on ONE persistent
    // WorkerPool and reuses the workers' output arrays from piece to piece: same rows as a one-shot parse, for several piece
    // sizes and pool sizes.  Built with -fsanitize=thread as well (argv[2] = "pool" runs only this part).
    {
        std::ifstream f2(argv[1], std::ios::binary);
        std::stringstream s2; s2 << f2.rdbuf();
        const std::string all = s2.str();
        jsonl::Trace want;
        jsonl::parse(all.data(), all.size(), 1, 0, 1, want);
        for (int threads : {2, 3, 8}) {
            jsonl::WorkerPool pool(threads);
            for (size_t piece : {(size_t)70000, (size_t)300000, all.size()}) {
                std::vector<jsonl::Trace> parts;  // reused across the pieces
                std::vector<int8_t> imv, mv;
                std::vector<uint16_t> ws;
                size_t pos = 0, lines = 0;
                uint32_t tau = 0;
                while (pos < all.size()) {
                    size_t use = std::min(piece, all.size() - pos);
                    if (pos + use < all.size()) {
                        size_t e = use;
                        while (e > 0 && all[pos + e - 1] != '\n') e--;
                        if (e == 0) { const size_t nl = all.find('\n', pos + use); e = (nl == std::string::npos ? all.size() : nl + 1) - pos; }
                        use = e;
                    }
                    uint32_t t_out = tau;
                    lines += jsonl::parse_parts(all.data() + pos, use, threads, tau, lines + 1, parts, t_out, &pool);
                    tau = t_out;
                    for (auto& t : parts) {
                        imv.insert(imv.end(), t.input_mv.begin(), t.input_mv.end());
                        mv.insert(mv.end(), t.mv.begin(), t.mv.end());
                        ws.insert(ws.end(), t.write_sym.begin(), t.write_sym.end());
                        t.clear_keep_capacity();
                    }
                    pos += use;
                }
                if (tau != want.tau || imv != want.input_mv || mv != want.mv || ws != want.write_sym) {
                    printf("MISMATCH threads %d piece %zu\n", threads, piece);
                    return 1;
                }
            }
        }
        printf("piecewise pool parse ok\n");
    }
}
