/* TEST INFRASTRUCTURE ONLY: empty stand-in for the htslib header of the same name, so that the reference's
 * pepper/modules/headers/dataio/bam_handler.h (type_read, CigarOp, CIGAR_OPERATIONS) can be parsed without htslib.
 * Only the three pointer types its BAM_handler class declaration mentions are declared; nothing is implemented. */
#ifndef PV_HTS_STUB_TYPES
#define PV_HTS_STUB_TYPES
struct htsFile; struct hts_idx_t; struct bam_hdr_t;
#endif
