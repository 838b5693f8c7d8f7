// stands in for Thirdparty/MB/ctmf.h (median filter of the dense MSA solver; out of scope)
