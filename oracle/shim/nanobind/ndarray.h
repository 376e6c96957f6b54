// oracle shim: see nanobind.h (test infrastructure only)
#pragma once
#include "nanobind.h"
