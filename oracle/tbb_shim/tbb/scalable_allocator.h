// empty stand-in: included (source/main.cpp:22) but unused by the reference. ORACLE BUILD ONLY.
#pragma once
