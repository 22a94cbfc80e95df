/* config.h for the BLAS-free oracle build of the UNMODIFIED reference sources:
 * ATRIP_USE_DGEMM left undefined selects the reference's own naive five-deep
 * loops (Equations.cxx:685-727), an independent second evaluation of Tijk. */
#ifndef ATRIP_ORACLE_CONFIG_H
#define ATRIP_ORACLE_CONFIG_H
#define ATRIP_DEBUG 1
/* ATRIP_NO_OUTPUT comes from the command line: Debug.hpp is reached before config.h in some units */
#endif
