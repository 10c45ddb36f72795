/* oracle/shim: stand-in for BamTools <api/BamReader.h> -- TEST INFRASTRUCTURE.
 * utils.h:42-44 includes the BamTools API only for CigarOp (utils.h:63). */
#ifndef ORC_SHIM_API_BAMREADER_H
#define ORC_SHIM_API_BAMREADER_H
#include <stdint.h>
#ifndef ORC_SHIM_CIGAROP
#define ORC_SHIM_CIGAROP
namespace BamTools { struct CigarOp { char Type; uint32_t Length; CigarOp(char t = 0, uint32_t l = 0) : Type(t), Length(l) {} }; }
#endif
#endif
