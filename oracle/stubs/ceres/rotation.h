#include "stub_ceres.h"
