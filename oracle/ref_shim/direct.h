// shim: <direct.h> (MSVC) -- only _chdir is used (Duke/utilities.cpp:438)
#pragma once
#include <unistd.h>
static inline int _chdir(const char *p) { return chdir(p); }
