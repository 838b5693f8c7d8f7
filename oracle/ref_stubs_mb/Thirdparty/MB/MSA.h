// forwards to the reference's REAL Thirdparty/MB/MSA.h for the `ref_mb` target of oracle/Makefile (the other reference
// builds resolve this include to the stub in ref_stubs/, because the dense solver is out of scope there)
#include REF_MSA_H
