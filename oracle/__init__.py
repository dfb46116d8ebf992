"""CPU oracle for the AvatarCLIP appearance-optimisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is shipped or measured as
the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / the timed CPU baseline.  The product path
(``avatarclip_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Contents (each function cites the reference file:line it restates; paths are
relative to ``/root/reference/AvatarGen/AppearanceGen``):

* ``oracle.neus``        -- positional encoding, SDF / colour / variance nets,
                            hierarchical sampling, NeuS compositing
                            (``models/embedder.py``, ``models/fields.py``,
                            ``models/renderer.py``); autograd supplies gradients.
* ``oracle.neus_manual`` -- the same forward *and an explicit hand-derived
                            backward* (no autograd), step for step the sequence
                            the CUDA kernels execute.  Checked against
                            ``oracle.neus`` autograd in the CPU test-suite.
* ``oracle.loss``        -- shading, canvas scatter and losses of
                            ``main.py:425-534``; camera / ray helpers of
                            ``models/dataset.py`` and ``models/utils.py``.
* ``oracle.clip_vit``    -- CLIP ViT-B/32 image tower (openai/CLIP, un-vendored
                            third-party dependency, unpinned in
                            ``requirements.txt:12``): published architecture
                            restated; **parity unpinned** (no reference test or
                            golden vector exists for it, weights not on disk).
* ``oracle.lbs``         -- ``my_lbs`` / ``batch_rodrigues``
                            (``models/utils.py:72-106,176-224``) with the
                            smplx helpers restated from their in-tree copies
                            (``drive.py:51-160``); **parity unpinned** beyond
                            the reference-import check (no SMPL model on disk).

Pinning status:
* the NeuS half is pinned against the *unmodified reference modules imported
  from /root/reference* (``oracle/pin_against_reference.py`` generates
  ``tests/golden/neus_*.pt`` from the reference itself);
* ``oracle.lbs`` against the reference's ``my_lbs`` run in place
  (``oracle/pin_lbs.py`` -> ``tests/golden/lbs_small.pt``);
* ``oracle.loss`` (loss stage main.py:417-534, schedules :571-586, ``lookat`` /
  ``sphere_coord``, ``gen_rays_pose`` / ``gen_rays_silhouettes`` /
  ``near_far_from_sphere``) against the reference's own SOURCE LINES, cut out of
  the files and executed in place (``oracle/pin_loss_stage.py`` ->
  ``tests/golden/loss_stage.pt``): main.py and models/{utils,dataset}.py do not
  import here, their arithmetic is plain torch / numpy;
* the CPU tests compare the restatement with those files on every run.
Still unpinned: the CLIP tower itself (third-party, weights not on disk).
"""
