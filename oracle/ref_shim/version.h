/* Stand-in for the header the reference's CMake generates from version.h.in
 * (reference: version.h.in:1-12, CMakeLists.txt:16-21).  Only main.cpp uses it. */
#pragma once
static const char *__g_prog_version = "0.5.0";
static const char *__g_build_type = "oracle";
