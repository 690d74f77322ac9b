"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU statement of the reference's algorithm for the prediction hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may import
anything from here, and only as the checker / the reported baseline -- never as the product.

Two kinds of content:
  * oracle/_ref/*.so  -- the reference's OWN C++ (stardist/lib/stardist2d.cpp, stardist3d.cpp,
    stardist3d_impl.cpp + vendored Clipper / Qhull / nanoflann) compiled unmodified from
    /root/reference by oracle/Makefile (setup.py:104-118 flags, gcc 13, -O2, OpenMP).  This is the
    ground truth for NMS 2D/3D and polyhedron_to_label.                      [parity: PINNED]
  * numpy / torch restatements of the Python-side logic whose modules cannot be imported here
    (stardist/nms.py, geometry/geom2d.py, models/base.py need csbdeep / scikit-image / TF, which
    are not installed): every function cites the reference lines it follows.
      - nms_np / sparse gather logic: pinned against the reference ext through golden vectors.
      - geom2d_np.polygon (skimage.draw.polygon, un-vendored, version unpinned): restated from the
        published skimage >= 0.18 rule; no reference run available.            [parity: UNPINNED]
      - unet_torch (Keras/TF U-Net via csbdeep): torch-CPU fp32 stand-in; no tensor-level golden
        exists in the reference.                                              [parity: UNPINNED]
"""
