// shim: <tchar.h> (MSVC) -- nothing from it is used on the compiled path
#pragma once
