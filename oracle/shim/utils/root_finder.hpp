// oracle/shim/utils/root_finder.hpp -- shadows back_end/include/utils/root_finder.hpp (complex eigenvalues, polynomial root
// isolation) for the reference-header pin build.  The only users are Piece::getMaxDotValueNorm / getMaxDDotValueNorm in
// se2traj.hpp, which nothing on the optimizeSE2Traj path calls ("root_finder.hpp (unreachable)"); the
// stubs below exist so those members compile and abort if ever reached.
#pragma once
#include <cstdlib>
#include <set>
#include <Eigen/Eigen>
namespace RootFinder {
inline Eigen::VectorXd polySqr(const Eigen::Ref &) { std::abort(); }
inline double polyVal(const Eigen::Ref &, double, bool = false) { std::abort(); }
inline std::set<double> solvePolynomial(const Eigen::Ref &, double, double, double, bool = true) { std::abort(); }
} // namespace RootFinder
