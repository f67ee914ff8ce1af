#pragma once
// stub: NormalDistributionsTransform.cpp includes the front end header but uses nothing from it
