// stands in for <Thirdparty/libelas/src/elas.h> (included by include/frame.h; ElasMatch is never called)
