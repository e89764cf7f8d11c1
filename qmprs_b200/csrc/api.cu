// Library identification for the loader (qmprs_b200/_lib.py).
#include "qmprs_b200.h"
extern "C" int qm_version(void) { return 100; }
