// oracle shim (test infrastructure only): F/misc.h:13-15 maps `foreach` to BOOST_FOREACH.
// boost is absent in this image; a C++11 range-for has the same semantics for the
// std containers the reference iterates over.
#pragma once
#define BOOST_FOREACH(decl, cont) for (decl : cont)
