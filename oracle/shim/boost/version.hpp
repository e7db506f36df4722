// Shim for compiling the reference's hot-path translation units without Boost
// (test infrastructure only; see oracle/README.md).  Claims a post-1.38 Boost
// so the reference picks <boost/spirit/include/classic.hpp>.
#pragma once
#define BOOST_VERSION 106000
