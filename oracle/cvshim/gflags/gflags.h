// oracle/cvshim/gflags/gflags.h — TEST INFRASTRUCTURE.  Empty stand-in: the
// reference's commfunc.h:16 includes gflags, but only main.cc (not compiled
// into oracle/_ref) uses it.
#pragma once
