// main.cpp -- the drop-in command line: TrimeshTracer <width> <height> <spp> <datafile>
// (main.cpp:248-345 of the reference).  Everything lives behind the C ABI in libtmpt.so.
#include "../../include/tmpt.h"

int main(int argc, const char** argv) { return tmpt_main(argc, argv); }
