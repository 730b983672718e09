"""tdnn-f_nas_b200: B200-native (sm_100a) training hot path of skhu101/TDNN-F_NAS.

The directory name contains a hyphen (it is the name the build contract asks for); import it
through the `tdnnf_nas_b200` alias package at the repository root:

    from tdnnf_nas_b200 import capi, nnet3
"""
__version__ = "0.1.0"
