// C-ABI glue (include/unet3d_b200.h): error string + model-level entry points.
#include <string>

#include "../../include/unet3d_b200.h"
#include "u3d.h"

namespace u3d {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const char* last_error() { return g_last_error.c_str(); }
}  // namespace u3d

extern "C" const char* unet3d_last_error(void) { return u3d::last_error(); }
