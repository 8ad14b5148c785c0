#include "cv_shim.h"
