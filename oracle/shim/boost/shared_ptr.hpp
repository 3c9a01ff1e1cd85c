// oracle shim (test infrastructure only): F/misc.h:17-19 does `using namespace boost`
// for shared_ptr and the pointer casts; std:: equivalents are drop-in.
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
}
