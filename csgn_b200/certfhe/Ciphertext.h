// Ciphertext.h -- forwarding header: the reference ships one header per class
// (src/Ciphertext.h); here all of them are declared in certFHE.h.
#include "certFHE.h"
