"""Python-level mirror of the reference's native extension modules (stardist.lib.stardist2d /
stardist.lib.stardist3d): same callable names, argument order and dtypes, backed by
libstardist_b200.so."""
