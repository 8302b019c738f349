"""oracle/ -- CPU restatement of the MTM hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under this directory is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker / the timed CPU
baseline -- never as a fallback for the CUDA path (the product raises when
``libmtm_b200.so`` or a GPU is missing).

What is restated, and from where
--------------------------------
The reference (MTM 2.0.1, ``/root/reference``) is ~490 lines of Python whose
arithmetic lives in third-party packages that are NOT vendored in the tree:

* ``opencv-python-headless >= 4.5.4`` (``setup.py:23``; 4.13.0 in this image):
  ``cv2.matchTemplate`` (call site ``MTM/__init__.py:92``), ``cv2.minMaxLoc``
  (``:226``), ``cv2.dnn.NMSBoxes`` (``MTM/NMS.py:78``).
* ``scikit-image`` (unpinned, ``setup.py:24``; ABSENT from this image):
  ``skimage.feature.peak_local_max`` (``MTM/__init__.py:45``).
* ``scipy`` (unpinned, ``setup.py:25``; 1.18.1 here): ``scipy.signal.find_peaks``
  (``MTM/__init__.py:34,40``).

Modules
-------
``ncc_exact``      exact-integer TM_CCOEFF_NORMED following OpenCV's published
                   ``common_matchTemplate`` epilogue (C via ctypes + numpy twin).
``peaks``          restated ``peak_local_max`` / ``find_peaks(height=)`` subset.
``nms_port``       restated ``cv2.dnn.NMSBoxes`` (NMSFast_ + jaccardDistance).
``mtm_port``       restated MTM orchestration (``MTM/__init__.py:22-296``,
                   ``MTM/NMS.py:20-84``) on top of live cv2 + ``peaks``.
``synth``          re-export of the repo-level ``workloads`` module (seeded synthetic inputs of
                   SURVEY.md section 8(d); the benches import ``workloads`` directly).
``ref_loader``     imports the UNMODIFIED reference from /root/reference (build
                   container only) to pin the restatement and make fixtures.

Pinning status: ``mtm_port``/``ncc_exact``/``nms_port`` are pinned against (a) the
reference itself run in the build container (tests/golden/*.json made by
``oracle/make_golden.py``), (b) the reference's own known answers
(Tutorial3-SpeedingUp.ipynb cells 10/14/21, ``MTM/NMS.py:86-96``), and (c) live
cv2 at test time.  ``peaks.peak_local_max`` restates an absent dependency from
its published algorithm; the reference holds no offline-reproducible golden
vector for the multi-object path, so that one function is "parity unpinned" by
reference-owned vectors (see DESIGN.md, section Oracle).
"""
