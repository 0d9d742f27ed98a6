// -*- C++ -*-
/*
 * Stand-in for <Rcpp.h> in the oracle/_ref build (TEST INFRASTRUCTURE; R and Rcpp are not in this
 * image).  /root/reference/src/Clusterer.h only *declares* two functions with Rcpp types
 * (Clusterer.h:40-41,57), so empty types suffice.  The header also pulls in what Rcpp.h pulls in
 * transitively (map, vector, ctime ...), and pins the reference's only nondeterminism:
 * Clusterer() and reseed() call srand(time(0) [+ offset]) (Clusterer.cpp:25, Clusterer.h:51-53);
 * time(x) is redirected to a settable value so a run can be repeated.
 */
#ifndef DPR_RCPP_SHIM_H
#define DPR_RCPP_SHIM_H
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <locale>
#include <map>
#include <string>
#include <vector>

namespace Rcpp {
class NumericMatrix {};
class List {};
}  // namespace Rcpp

extern "C" long dpr_ref_fixed_time(void);
#define time(x) dpr_ref_fixed_time()
#endif
