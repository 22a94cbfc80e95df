/* config.h for the oracle build of the UNMODIFIED reference sources: the
 * switches autoconf would have written (reference configure.ac:88-147).
 * dgemm path (Equations.cxx:492-684), rank-0 logging silenced (Debug.hpp:88-93). */
#ifndef ATRIP_ORACLE_CONFIG_H
#define ATRIP_ORACLE_CONFIG_H
#define ATRIP_USE_DGEMM 1
#define ATRIP_DEBUG 1
/* ATRIP_NO_OUTPUT comes from the command line: Debug.hpp is reached before config.h in some units */
#endif
