/* Forwarding header: lets a program written against the reference's own file names (#include "util.h") build against
 * the B200 engine.  Everything lives in ../spmv_fpga_compat.h. */
#include "../spmv_fpga_compat.h"
