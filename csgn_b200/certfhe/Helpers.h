// Helpers.h -- forwarding header: the reference ships one header per class
// (src/Helpers.h); here all of them are declared in certFHE.h.
#include "certFHE.h"
