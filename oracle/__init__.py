"""CPU oracle for the DATMO optical-flow hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker or the CPU baseline,
never as the thing shipped: the product package
(``datmo_using_optical_flow_b200``) never imports ``oracle`` and raises when
its CUDA library is missing.

What it restates (reference = /root/reference/Optical_flow/main.py, pure
Python over third-party wheels; nothing in it is compiled, so there is no
``oracle/_ref`` binary to build):

  module              restates                                        pinned by
  ------------------  ----------------------------------------------  -----------------------------------------
  bev_np              compute_bev_grid / filter_points_in_roi /       reference main.py imported here (stubbed
                      increase_point_density  (main.py:30-57,98-126)  open3d/matplotlib/shapely) -> tests/golden
  farneback_np        cv2.calcOpticalFlowFarneback as called at       live cv2 4.13.0 (opencv-python-headless,
                      main.py:132-142 (OpenCV video/optflowgf.cpp,    unpinned by the reference) + golden flow
                      source not on disk; published algorithm)        vectors produced through main.py
  masks_np            continuity_mask + inline moving-cell filter     reference main.py -> tests/golden
                      (main.py:224-228, 596-609)
  dbscan_np           sklearn.cluster.DBSCAN as called at             live sklearn 1.9.0 + golden labels
                      main.py:231-259, restated as the grid rule      produced through main.py
  cluster_np          extract_cluster_data (main.py:402-434)          reference main.py -> tests/golden
  ransac_np           Open3D segment_plane as called at main.py:73    PARITY UNPINNED: open3d is absent from
                                                                      this image and not vendored; restated
                                                                      from the published algorithm only
  reference_port      the stage functions of main.py, calling cv2 /   is the reference's own call sequence;
                      sklearn exactly as main.py does (the CPU        used as cpu_baseline kind="port"
                      baseline that can travel to the GPU box)

Third-party dependencies holding the arithmetic (none pinned by the
reference: README.md:16-28 lists bare names, there is no requirements file):
opencv-python (4.13.0.92 in this image), scikit-learn (1.9.0), open3d (absent).
"""
