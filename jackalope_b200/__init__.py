"""jackalope_b200 -- B200-native (CUDA sm_100a) Illumina read generation with
jackalope's ``illumina()`` argument surface.  See DESIGN.md."""
from .genome import Haplotypes, HapChromMuts, RefGenome, create_genome, random_genome, random_haplotypes  # noqa: F401
from .illumina import Context, check_illumina_args, default_context, illumina  # noqa: F401
from .pacbio import check_pacbio_args, pacbio, sample_read_plan  # noqa: F401
from .profiles import (JackalopeError, builtin_illumina_profiles, find_profile_file, flatten_profile,  # noqa: F401
                       format_profile, read_profile, seq_sys_by_read_length)

__all__ = ["illumina", "pacbio", "create_genome", "Context", "RefGenome", "Haplotypes", "HapChromMuts", "random_genome", "random_haplotypes",
           "read_profile", "flatten_profile", "JackalopeError"]
