"""ambivert_b200 -- B200 (sm_100a) backend for AmBiVErT's alignment hot path.

  ambivert_b200.align   drop-in for the reference's `ambivert.align` package (same names, same
                        ctypes structs, same call signatures), bound to align_c.b200.so
  ambivert_b200.batch   numpy-facing wrappers of the batched C ABI (include/ambivert_align.h part 2)
  ambivert_b200.host    the pipeline stages that sit on the path (merge_overlaps,
                        match_to_reference, align_to_reference) calling the batched ABI once per stage
  ambivert_b200.dist    query sharding over the GPUs of one box + the single NCCL gather

There is no CPU implementation of any alignment in this package: importing `align`/`batch` without
the built CUDA library raises, and every call fails loudly without a GPU.
"""
__version__ = "0.1.0"
