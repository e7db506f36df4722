// Shim: common/bpmatrix.h only needs the name options_description (by
// reference, in a member that the oracle never defines or calls) and, through
// the real header's transitive includes, boost::shared_ptr.
#pragma once
#include "shared_ptr.hpp"
namespace boost { namespace program_options { class options_description; } }
