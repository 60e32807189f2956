/* Forwarding header: the reference's include path <lambda_snark/utils.h>
 * (cpp-core/include/lambda_snark/utils.h) resolves to the B200 C ABI. */
#pragma once
#include "../lambda_snark_b200.h"
