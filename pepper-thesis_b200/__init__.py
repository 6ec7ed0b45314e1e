"""pepper-thesis_b200: B200-native (sm_100a) pileup-summary + recurrent-inference hot path of PEPPER r0.8.

Sub-modules (imported lazily; nothing here touches CUDA at import time):
  nativebuild  compile every native artefact in-tree (nvcc / gcc / g++)
  read_batch   packed SoA read batch (host side of include/pepper_b200.h PvReadBatch)
  synth        seeded synthetic pileups
  capi         ctypes binding of the C-ABI library libpepper_b200.so
  summarizer   drop-in for pepper_variant AlignmentSummarizer.create_summary
  models       TransducerGRU contracts (variant LSTM model M-A, polisher GRU model M-B) on the CUDA path
  pipeline     region-sharded summary -> inference driver (1..8 GPUs)
  candidate_filter / vcf_writer   stage 3: device candidate filter, pysam-free VCFWriter (BGZF .vcf.gz)
"""
__version__ = "0.1.0"
