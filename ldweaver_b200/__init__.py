"""ldweaver_b200 -- B200-native drop-in for LDWeaver's genome-wide pairwise-LD hot path.

Only what the path needs lives here: ``csrc/`` (hand-written sm_100a CUDA kernels + the C ABI of
``include/ldw.h``) and the host-side mirror of the reference's R interface (``api``)."""
from .api import (CdsVar, MIPlan, MIScanResult, make_blocks, partition_blocks, perform_MI_computation,  # noqa: F401
                  SnpDat, acgtn2num, estimate_Hamming_distance_weights, parse_fasta_alignment,  # noqa: F401
                  parse_fasta_SNP_alignment, snp_dat_from_alignment_matrix, snp_dat_from_codes, mergeNsort_sr_links,  # noqa: F401
                  runARACNE, finish_sr_links, write_sr_tsv, write_lr_tsv, SrLinks, read_LongRangeLinks,  # noqa: F401
                  read_ShortRangeLinks, analyse_long_range_links, DeviceGroup)  # noqa: F401

__all__ = ["CdsVar", "MIPlan", "MIScanResult", "make_blocks", "partition_blocks", "perform_MI_computation", "SnpDat", "acgtn2num", "estimate_Hamming_distance_weights", "parse_fasta_alignment",
           "parse_fasta_SNP_alignment", "snp_dat_from_alignment_matrix", "snp_dat_from_codes", "mergeNsort_sr_links", "runARACNE", "finish_sr_links",
           "write_sr_tsv", "write_lr_tsv", "SrLinks", "read_LongRangeLinks", "read_ShortRangeLinks", "analyse_long_range_links", "DeviceGroup"]
