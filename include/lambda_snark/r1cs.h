/* Forwarding header: the reference's include path <lambda_snark/r1cs.h>
 * (cpp-core/include/lambda_snark/r1cs.h) resolves to the B200 C ABI. */
#pragma once
#include "../lambda_snark_b200.h"
