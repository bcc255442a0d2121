// utils.h -- what the reference's src/utils.h gives every translation unit: the BIT
// macro and the standard headers user code relies on.  `using namespace std` in a
// header is not something we would choose, but programs written against the reference
// (its tests/*.cpp among them) use cout/endl/string unqualified after including
// certFHE.h, so the drop-in keeps that promise.
#ifndef CSGN_CERTFHE_UTILS_H_
#define CSGN_CERTFHE_UTILS_H_

#define BIT(X) X & 0x01

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <bitset>
#include <chrono>
#include <iostream>
#include <string>
#include <vector>

using namespace std;

#endif
