#ifndef SHIM_ASSERTION_UTILITIES_H_
#define SHIM_ASSERTION_UTILITIES_H_
#include <cmath>
#include <sstream>
#include <string>
#include "openmm/OpenMMException.h"
namespace OpenMM {
inline void throwException(const char* file, int line, const std::string& details) {
    std::stringstream m;
    m << "Assertion failure at " << file << ":" << line;
    if (!details.empty()) m << ".  " << details;
    throw OpenMMException(m.str());
}
}  // namespace OpenMM
#define ASSERT(cond) {if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, "");};
#define ASSERT_EQUAL(expected, found) {if (!((expected) == (found))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_EQUAL_TOL(expected, found, tol) {double _scale_ = std::abs(expected) > 1.0 ? std::abs(expected) : 1.0; if (!(std::abs((expected)-(found))/_scale_ <= (tol))) {std::stringstream details; details << "Expected "<<(expected)<<", found "<<(found); OpenMM::throwException(__FILE__, __LINE__, details.str());}};
#define ASSERT_USUALLY_EQUAL_TOL(expected, found, tol) ASSERT_EQUAL_TOL(expected, found, tol)
#define ASSERT_VALID_INDEX(index, vector) {if (index < 0 || index >= (int) vector.size()) OpenMM::throwException(__FILE__, __LINE__, "Index out of range");};
#endif
