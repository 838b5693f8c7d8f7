// stands in for <sophus/se3.h> (included by include/pnpmatch.h, nothing of it is used on this path)
