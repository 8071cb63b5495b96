// Shim header (test infrastructure): everything lives in core/core.hpp.
#include "../core/core.hpp"
