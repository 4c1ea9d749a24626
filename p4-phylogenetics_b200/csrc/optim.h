// optim.h -- derivative-free minimisers of csrc/praxis.cpp (host code; the objective runs on the GPU).
#pragma once
#include <functional>

namespace p4b {

// The objective may overwrite its argument (the reference's p4_unWindParameters writes clamped values back into
// the vector it is handed, Pf/p4_treeOpt.c:568-571).
typedef std::function<double(double *)> Objective;

// Brent's principal-axis method (Pf/brent.c praxis): minimise f from x; returns the minimum found, x at it.
class Praxis {
  public:
    struct State;
    explicit Praxis(int n);
    ~Praxis();
    Praxis(const Praxis &) = delete;
    Praxis &operator=(const Praxis &) = delete;
    double minimize(double tol, double h, double *x, Objective f);

  private:
    State *S;
};

// Powell's direction-set method inside the box [lo, hi]; stops when an iteration improves f by less than ftol
// (relative) or after maxEvals evaluations.  Returns the minimum found, x at it; *nEvals is incremented.
double boundedPowell(int n, double *x, const double *lo, const double *hi, Objective f, double xtol, double ftol, long maxEvals, long *nEvals);

}  // namespace p4b
