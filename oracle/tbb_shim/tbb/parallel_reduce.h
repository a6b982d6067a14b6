// empty stand-in: included (source/scene.cpp:4) but unused by the reference. ORACLE BUILD ONLY.
#pragma once
