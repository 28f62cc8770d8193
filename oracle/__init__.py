"""CPU oracle for the PEANUT perception hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``peanut_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may.

Each module restates one stage of the reference in plain PyTorch / numpy and cites the reference
file:line it follows.  Pinning status:

* ``oracle.mapper``   - pinned against the reference's own ``nav/agent/mapping.py`` executed in the build
  container (fixtures in tests/golden/semmap_*.npz, generator tests/golden/make_semmap_golden.py).
* ``oracle.prednet``  - pinned against the reference's own model files (resnet.py, res_layer.py, psp_head.py,
  decode_head.py, encoder_decoder.py, wrappers.py built from nav/pred_model_cfg.py) executed in the build
  container over a restated mmcv-1.6.0 shim: bit-identical logits and stage outputs are required before
  tests/golden/prednet_*.npz are written (generator tests/golden/make_prednet_golden.py, shim mmcv_shim.py).
* ``oracle.preproc`` / ``oracle.agent_prediction`` / ``oracle.map_state`` / ``oracle.goal_map`` - pinned against the
  UNMODIFIED methods of nav/agent/agent_helper.py and nav/agent/agent_state.py, cut out with ``ast`` and executed on stub
  objects (generators tests/golden/make_{preproc,update_prediction,map_state,goal_map}_golden.py).
* ``oracle.maskrcnn`` - PARITY UNPINNED: detectron2 0.6 is absent; restated from the config
  nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml and detectron2's published semantics.
"""
