#ifndef NW_REF_SHADOW_LINEARSOLVERTYPES_H
#define NW_REF_SHADOW_LINEARSOLVERTYPES_H
#include <KokkosInterface.h>
#endif
