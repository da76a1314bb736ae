// oracle/shim/ompl/base/spaces/DubinsStateSpace.h -- stand-in for OMPL (ros-noetic-ompl, absent here) with exactly the surface
// front_end/include/front_end/kino_astar.h:246-266 uses: StateSpacePtr, DubinsStateSpace(radius), ScopedState<> with operator[],
// operator() and reals(), StateSpace::distance and ::interpolate.  The curve itself is uneven_planner_b200/csrc/dubins.h, the
// restatement of OMPL's published algorithm that the product uses too: the pin built on this shim covers KinoAstar's own logic
// (search, costs, collision checks, path assembly), not OMPL's arithmetic.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
#include <vector>
#include "../../../../../uneven_planner_b200/csrc/dubins.h"
namespace ompl {
namespace base {
struct State { double v[3] = {0, 0, 0}; };
class StateSpace {
public:
    virtual ~StateSpace() {}
    virtual double distance(const State *a, const State *b) const = 0;
    virtual void interpolate(const State *from, const State *to, double t, State *out) const = 0;
};
typedef std::shared_ptr<StateSpace> StateSpacePtr;
class DubinsStateSpace : public StateSpace {
public:
    explicit DubinsStateSpace(double turningRadius = 1.0, bool = false) : rho_(turningRadius) {}
    double distance(const State *a, const State *b) const override { return ualm_dubins::distance(a->v, b->v, rho_); }
    void interpolate(const State *from, const State *to, double t, State *out) const override
    {
        const ualm_dubins::Path p = ualm_dubins::shortest(from->v, to->v, rho_);
        ualm_dubins::interpolate(from->v, p, rho_, t, out->v);
    }
private:
    double rho_;
};
template <class T = StateSpace>
class ScopedState {
public:
    explicit ScopedState(const StateSpacePtr &) {}
    double &operator[](int i) { return s_.v[i]; }
    State *operator()() { return &s_; }
    std::vector<double> reals() const { return std::vector<double>(s_.v, s_.v + 3); }
private:
    State s_;
};
} // namespace base
} // namespace ompl
