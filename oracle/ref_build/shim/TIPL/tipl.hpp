// TEST INFRASTRUCTURE ONLY (oracle build).  Not product code.
//
// Minimal stand-in for the un-vendored TIPL library (frankyeh/TIPL, cloned at HEAD by the
// reference's own build: /root/reference/CMakeLists.txt:83), providing ONLY the symbols that
// /root/reference/unet.hpp and /root/reference/unet.cpp touch, so that those two files
// compile UNCHANGED, in place, against the image's libtorch:
//   tipl::split, tipl::split_by_line_breaks   (unet.cpp:27,115,120)
//   tipl::par_for                             (unet.cpp:230)
//   tipl::progress                            (unet.cpp:248)
//   tipl::vector<3>, tipl::shape<3>           (unet.hpp:37-38; printed at unet.cpp:283)
// The only semantic content is string splitting (drop '\r', drop empty lines).
#pragma once
#include <array>
#include <cstddef>
#include <initializer_list>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace tipl {

template <int N, class T = float>
struct vector {
    std::array<T, N> v{};
    vector() = default;
    vector(std::initializer_list<T> l) {
        int i = 0;
        for (auto e : l) if (i < N) v[i++] = e;
    }
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
    const T* begin() const { return v.data(); }
    const T* end() const { return v.data() + N; }
};
template <int N, class T>
std::ostream& operator<<(std::ostream& o, const vector<N, T>& r) {
    for (int i = 0; i < N; ++i) o << (i ? " " : "") << r[i];
    return o;
}

template <int N>
struct shape {
    std::array<unsigned int, N> v{};
    shape() = default;
    shape(std::initializer_list<unsigned int> l) {
        int i = 0;
        for (auto e : l) if (i < N) v[i++] = e;
    }
    unsigned int& operator[](size_t i) { return v[i]; }
    const unsigned int& operator[](size_t i) const { return v[i]; }
    size_t size() const {
        size_t s = 1;
        for (auto e : v) s *= e;
        return s;
    }
};
template <int N>
std::ostream& operator<<(std::ostream& o, const shape<N>& r) {
    for (int i = 0; i < N; ++i) o << (i ? " " : "") << r[i];
    return o;
}

inline std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream in(s);
    while (std::getline(in, cur, sep)) out.push_back(cur);
    return out;
}

inline std::vector<std::string> split_by_line_breaks(const std::string& s) {
    std::vector<std::string> out;
    std::string cur;
    std::istringstream in(s);
    while (std::getline(in, cur, '\n')) {
        while (!cur.empty() && (cur.back() == '\r' || cur.back() == ' ')) cur.pop_back();
        if (!cur.empty()) out.push_back(cur);
    }
    return out;
}

template <class F>
void par_for(size_t n, F&& f) {
    for (size_t i = 0; i < n; ++i) f(i);
}

struct progress {
    explicit progress(const char*) {}
};

}  // namespace tipl
