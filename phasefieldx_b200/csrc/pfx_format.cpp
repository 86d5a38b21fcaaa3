// pfx_format.cpp — text formatting for the ASCII VTU output.
//
// The reference hands its fields to dolfinx's VTKFile writer every load step; on a 10^6-node
// mesh the ASCII conversion of 4 doubles per node is what the step waits for once the solves
// run on the GPU.  std::to_chars writes the shortest text that parses back to the same double
// (so the file is bit-exact on read-back) at tens of nanoseconds per number; rows are
// formatted by a few host threads into disjoint slices of the caller's buffer.
#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/pfx_b200.h"

namespace {

constexpr int kMaxF64 = 25;   // "-1.2345678901234567e-308" + separator
constexpr int kMaxI64 = 21;

template <class Row>
int64_t format_rows(int64_t n_rows, int64_t row_cap, char* out, int64_t cap, Row row) {
    if (n_rows <= 0) return 0;
    if (cap < n_rows * row_cap) return -1;
    int nt = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    nt = (int)std::min<int64_t>(nt, (n_rows + 65535) / 65536);
    if (nt <= 1) {
        char* p = out;
        for (int64_t i = 0; i < n_rows; ++i) p = row(i, p);
        return p - out;
    }
    // each thread formats its rows at the worst-case offset, then the slices are compacted
    std::vector<int64_t> len(nt);
    std::vector<std::thread> th;
    int64_t per = (n_rows + nt - 1) / nt;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            int64_t a = t * per, b = std::min(n_rows, a + per);
            char* p0 = out + a * row_cap;
            char* p = p0;
            for (int64_t i = a; i < b; ++i) p = row(i, p);
            len[t] = p - p0;
        });
    for (auto& t : th) t.join();
    int64_t pos = len[0];
    for (int t = 1; t < nt; ++t) {
        std::memmove(out + pos, out + (int64_t)t * per * row_cap, (size_t)len[t]);
        pos += len[t];
    }
    return pos;
}

}  // namespace

extern "C" {

int64_t pfx_format_f64(const double* v, int64_t n_rows, int32_t ncomp_in, int32_t ncomp_out, char* out,
                       int64_t cap) {
    if (!v || !out || ncomp_in < 1 || ncomp_out < ncomp_in) return -1;
    return format_rows(n_rows, (int64_t)ncomp_out * kMaxF64, out, cap, [=](int64_t i, char* p) {
        for (int c = 0; c < ncomp_out; ++c) {
            if (c < ncomp_in) {
                p = std::to_chars(p, p + kMaxF64, v[i * ncomp_in + c]).ptr;
            } else {
                *p++ = '0';
            }
            *p++ = ' ';
        }
        return p;
    });
}

int64_t pfx_format_i64(const int64_t* v, int64_t n, char* out, int64_t cap) {
    if (!v || !out) return -1;
    return format_rows(n, kMaxI64, out, cap, [=](int64_t i, char* p) {
        p = std::to_chars(p, p + kMaxI64, v[i]).ptr;
        *p++ = ' ';
        return p;
    });
}

}  // extern "C"
