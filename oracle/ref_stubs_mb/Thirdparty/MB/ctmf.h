#include REF_CTMF_H
