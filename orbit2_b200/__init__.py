"""orbit2_b200 -- B200-native (sm_100a) kernels + drop-in host module for ORBIT-2's Reslim hot path."""
__version__ = "0.1.0"
