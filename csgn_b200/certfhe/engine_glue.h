// engine_glue.h -- internal: how the certFHE classes reach the C ABI (include/csgn.h).
#ifndef CSGN_CERTFHE_ENGINE_GLUE_H_
#define CSGN_CERTFHE_ENGINE_GLUE_H_

#include "csgn.h"

#include <string>

namespace certFHE {
namespace glue {

// csgn_init on first use (Library::initializeLibrary calls it eagerly).  Throws
// certFHE::Error when no sm_100 device is usable: there is no CPU path to fall back to.
void ensure_engine();

// The job's peer communicator (Library::connectPeers), or null in a single-process run.
csgn_comm *comm();

// Turns a non-zero csgn_status into a certFHE::Error carrying csgn_last_error().
void check(int status, const char *what);

}  // namespace glue
}  // namespace certFHE

#endif
