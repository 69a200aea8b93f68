"""CPU oracle for the Exposure hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``exposure_b200``) never imports it and has no CPU fallback.

PARITY: PINNED TO THE REFERENCE'S PYTHON, NOT TO TENSORFLOW BINARIES.  The reference
(yuanming-hu/exposure @ 7bb838a) ships no golden vectors / known-answer tests and
TensorFlow 1.6 cannot be installed in this image.  What can run here is the reference's
own Python: tests/golden/make_reference_golden.py imports filters.py, agent.py,
critics.py, pdf_sample_layer.py, util.py, config_example.py, replay_memory.py and net.py
(GAN.__init__'s graph construction, run eagerly on pre-fed placeholders) from /root/reference
unmodified and executes them over tests/golden/tf1_shim (an eager stand-in for the TF-1
API subset they call, torch CPU underneath), with the shipped pretrained checkpoint and
with name-seeded weights; the outputs are committed as tests/golden/reference_golden.npz
and tests/test_reference_golden.py holds this oracle to them (<= 5e-6 in fp64, bounded by
fp32 rounding of constants): every Filter subclass forward / gradients, masked apply,
5-step agent_generator rollouts, critic / value, the generator / value / critic losses
and their gradients including the WGAN-GP double backward (the same script also records the
reference's cv2 debugger canvases, its thumbnails of the sample TIFFs and the draw sequences
of its ReplayMemory, which pin the host-side mirrors in exposure_b200/).  So op order, constants,
broadcasting, variable scopes and formulas are pinned to the reference source.  What
stays restated from TF 1.6's published kernels, in the shim exactly as here: exp / pow /
cos / tanh / sigmoid, clip_by_value / maximum / minimum tie rules, RGBToHSV / HSVToRGB
(colorspace_op.h), nn.moments, SAME convolution, floor(keep+U)/keep dropout, Adam.
Self-consistency checks (fp32-vs-fp64, finite differences, invariants) are in
tests/test_oracle_*.py.
"""
