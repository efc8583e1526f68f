// ORACLE SCAFFOLDING (test infrastructure): the reference is built with
// -DAGAST_GLOG so its own assert-based CHECK macros (agast/glog.h) are used;
// this header only has to exist for stray <glog/logging.h> includes.
#pragma once
#ifndef AGAST_GLOG
#define AGAST_GLOG
#endif
#include <agast/glog.h>
