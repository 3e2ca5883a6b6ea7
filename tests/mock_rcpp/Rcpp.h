// Rcpp.h -- TEST STAND-IN, not Rcpp.  R and Rcpp are not installed in the image this repository is built in;
// this header implements just the slice of the Rcpp API that glue/ehmm_glue.cpp uses, with plain C++
// containers, so that the glue can be compiled, linked against libehmm_b200.so and run by
// tests/glue_harness.cpp.  Semantics follow Rcpp where the glue depends on them: vectors and matrices are
// reference-counted handles (copying a NumericVector aliases it, clone() copies), matrices are column-major,
// List elements hold any of the value types, Environment::global_env() is one process-wide table (it holds
// .Random.seed), stop() throws.
#pragma once
#include <any>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef std::ptrdiff_t R_xlen_t;

namespace Rcpp {

[[noreturn]] inline void stop(const std::string &m) { throw std::runtime_error(m); }

class Dimension {
public:
    Dimension(int a, int b) : d{a, b} {}
    int d[2];
};

template <typename T> class Vector {
public:
    Vector() : p(std::make_shared<std::vector<T>>()) {}
    explicit Vector(int n) : p(std::make_shared<std::vector<T>>((size_t)n)) {}
    Vector(std::initializer_list<T> l) : p(std::make_shared<std::vector<T>>(l)) {}
    T *begin() { return p->data(); }
    const T *begin() const { return p->data(); }
    R_xlen_t size() const { return (R_xlen_t)p->size(); }
    T &operator[](R_xlen_t i) { return (*p)[(size_t)i]; }
    const T &operator[](R_xlen_t i) const { return (*p)[(size_t)i]; }
    std::shared_ptr<std::vector<T>> p;
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Vector<int> LogicalVector;

template <typename T> Vector<T> clone(const Vector<T> &v) {
    Vector<T> r;
    *r.p = *v.p;
    return r;
}

template <typename T> class Matrix {
public:
    Matrix() : r(0), c(0), p(std::make_shared<std::vector<T>>()) {}
    Matrix(int rows, int cols) : r(rows), c(cols), p(std::make_shared<std::vector<T>>((size_t)rows * cols)) {}
    int nrow() const { return r; }
    int ncol() const { return c; }
    T *begin() { return p->data(); }
    const T *begin() const { return p->data(); }
    T &operator()(int i, int j) { return (*p)[(size_t)i + (size_t)j * r]; }
    const T &operator()(int i, int j) const { return (*p)[(size_t)i + (size_t)j * r]; }
    int r, c;
    std::shared_ptr<std::vector<T>> p;
};
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;

class StringVector {
public:
    StringVector() {}
    StringVector(std::initializer_list<std::string> l) : v(l) {}
    const std::string &operator[](R_xlen_t i) const { return v[(size_t)i]; }
    R_xlen_t size() const { return (R_xlen_t)v.size(); }
    std::vector<std::string> v;
};
template <typename T> T as(const std::string &s) { return T(s); }

// NULL or a value (R's NULL default arguments)
template <typename T> class Nullable {
public:
    Nullable() : set(false) {}
    Nullable(const T &x) : set(true), v(x) {}
    bool isNotNull() const { return set; }
    bool isNull() const { return !set; }
    operator T() const { return v; }
    bool set;
    T v;
};

struct NamedValue {
    std::string name;
    std::any value;
};
class Named {
public:
    explicit Named(const char *n) : name(n) {}
    template <typename T> NamedValue operator=(const T &x) const { return NamedValue{name, std::any(x)}; }
    std::string name;
};

class List {
public:
    struct Attr {
        List *l;
        std::string key;
        void operator=(const Dimension &d) { l->dim[0] = d.d[0]; l->dim[1] = d.d[1]; }
    };
    List() {}
    explicit List(int n) : v((size_t)n), names((size_t)n) {}
    template <typename... A> static List create(const A &...a) {
        List l;
        (l.push(a), ...);
        return l;
    }
    std::any &operator[](R_xlen_t i) { return v[(size_t)i]; }
    const std::any &operator[](const std::string &n) const {
        for (size_t i = 0; i < names.size(); i++)
            if (names[i] == n) return v[i];
        throw std::runtime_error("no list element " + n);
    }
    template <typename T> T get(const std::string &n) const { return std::any_cast<T>((*this)[n]); }
    template <typename T> T get(R_xlen_t i) const { return std::any_cast<T>(v[(size_t)i]); }
    R_xlen_t size() const { return (R_xlen_t)v.size(); }
    Attr attr(const char *k) { return Attr{this, k}; }
    void push(const NamedValue &nv) { v.push_back(nv.value); names.push_back(nv.name); }
    std::vector<std::any> v;
    std::vector<std::string> names;
    int dim[2] = {0, 0};
};

class Environment {
public:
    struct Binding {
        std::any *slot;
        template <typename T> operator T() const {
            if (!slot->has_value()) throw std::runtime_error("object not found");
            return std::any_cast<T>(*slot);
        }
        template <typename T> void operator=(const T &x) { *slot = std::any(x); }
    };
    static Environment global_env() { return Environment(); }
    Binding operator[](const std::string &n) { return Binding{&table()[n]}; }
    static std::map<std::string, std::any> &table() {
        static std::map<std::string, std::any> t;
        return t;
    }
};

class RNGScope {      // GetRNGstate / PutRNGstate: the mock generator below has no cached state to sync
public:
    RNGScope() {}
    ~RNGScope() {}
};

// a deterministic stand-in for R's unif_rand() in (0, 1): the harness only needs draws that it can replay
inline uint64_t &mock_rng_state() {
    static uint64_t s = 0x9E3779B97F4A7C15ull;
    return s;
}
inline double unif_rand() {
    uint64_t &s = mock_rng_state();
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return ((double)(s >> 11) + 0.5) / 9007199254740992.0;
}

}  // namespace Rcpp

namespace R {
// R::rbinom(1, p) by inversion on one uniform, none drawn for p = 0 or 1 (nmath/rbinom.c, n = 1)
inline double rbinom(double n, double pp) {
    (void)n;
    if (pp == 0.0) return 0.0;
    if (pp == 1.0) return 1.0;
    const double pm = pp < 1.0 - pp ? pp : 1.0 - pp, q = 1.0 - pm, u = Rcpp::unif_rand();
    int ix = (u < q) ? 0 : 1;
    if (pp > 0.5) ix = 1 - ix;
    return (double)ix;
}
}  // namespace R
