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
int main(int argc, char** argv) {
    std::ifstream f(argv[1], std::ios::binary);
    std::stringstream ss; ss << f.rdbuf();
    std::string s = ss.str();
    // keep the first 8 lines; exact-size heap buffers so that ASan sees any read past the end
    size_t pos = 0; for (int i = 0; i < 8; i++) pos = s.find('\n', pos) + 1;
    s.resize(pos);
    size_t ok = 0, bad = 0;
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
}
