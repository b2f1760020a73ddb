"""Drop-in `pytorch3d` package (seam #3): only `pytorch3d.ops`, backed by
libdpm_b200.so.  Put `deeppointmap_b200/compat` on sys.path and the unmodified reference's
`-t3d` branches (network/encoder/utils.py:11-14,29-38,134-143; dataloader/transforms.py:10-14;
system/modules/utils.py:9-13) run on the B200 kernels."""
__version__ = "0.7.4+dpm_b200"
