"""kindel_b200 -- B200-native pileup/consensus engine behind kindel's API."""
__version__ = "1.2.1"

