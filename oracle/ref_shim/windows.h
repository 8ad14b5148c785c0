// shim: <windows.h> -- only what Utilities::folderScan (Duke/utilities.cpp:436-462) names, as inert stubs.
#pragma once
#include <wchar.h>
typedef void *HANDLE;
#define INVALID_HANDLE_VALUE ((HANDLE)(long)-1)
struct WIN32_FIND_DATA { wchar_t cFileName[260]; };
static inline HANDLE FindFirstFile(const wchar_t *, WIN32_FIND_DATA *) { return INVALID_HANDLE_VALUE; }
static inline int FindNextFile(HANDLE, WIN32_FIND_DATA *) { return 0; }
static inline int lstrlen(const wchar_t *s) { return (int)wcslen(s); }
