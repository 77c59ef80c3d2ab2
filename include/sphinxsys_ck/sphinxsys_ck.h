// sphinxsys_ck.h — umbrella header of the C++ host layer over libsphb200.so (see base.h for scope and conventions).
#ifndef SPHINXSYS_CK_H
#define SPHINXSYS_CK_H
#include "base.h"
#include "geometry.h"
#include "particles.h"
#include "configuration.h"
#include "fluid_dynamics.h"
#include "io_ck.h"
#endif
