// Compile-check stub of the handful of Rcpp / R API names rpkg/src/InterfaceUtils.cpp uses (R/Rcpp are not in this
// image).  It only has to make `g++ -fsyntax-only` meaningful; it is not a runtime.
#pragma once
#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
typedef struct SEXPREC* SEXP;
static SEXP R_NilValue = 0;
typedef void (*R_CFinalizer_t)(SEXP);
#ifndef TRUE
#define TRUE 1
#endif
inline SEXP R_MakeExternalPtr(void*, SEXP, SEXP) { return 0; }
inline void* R_ExternalPtrAddr(SEXP) { return 0; }
inline void R_ClearExternalPtr(SEXP) {}
inline void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, int) {}
inline SEXP PROTECT(SEXP s) { return s; }
inline void UNPROTECT(int) {}
namespace Rcpp {
template <typename T> struct Vec {
    std::vector<T> v;
    Vec() {}
    Vec(long n) : v((size_t)n) {}
    long size() const { return (long)v.size(); }
    T& operator[](long i) { return v[(size_t)i]; }
    T* begin() { return v.data(); }
};
typedef Vec<int> IntegerVector;
typedef Vec<double> NumericVector;
struct NumericMatrix : Vec<double> {
    int r, c;
    NumericMatrix(int rr, int cc) : Vec<double>((long)rr * cc), r(rr), c(cc) {}
    int nrow() const { return r; }
    int ncol() const { return c; }
};
struct Any {
    Any() {}
    template <typename T> Any(const T&) {}
};
struct List {
    std::map<std::string, Any> named;
    std::vector<Any> pos;
    List() {}
    List(size_t n) : pos(n) {}
    size_t size() const { return pos.size(); }
    Any& operator[](const char* k) { return named[k]; }
    Any& operator[](size_t i) { return pos[i]; }
    Any& operator[](int i) { return pos[(size_t)i]; }
};
template <typename T> T clone(const T& x) { return x; }
template <typename T> T as(const Any&) { return T(0, 0); }
inline void stop(const std::string& m) { throw std::runtime_error(m); }
}  // namespace Rcpp
inline double* REAL(Rcpp::Vec<double>& x) { return x.v.data(); }
inline int* INTEGER(Rcpp::Vec<int>& x) { return x.v.data(); }
