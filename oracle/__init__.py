"""CPU/fp32 oracle for the MADM diffusion feature-extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``madm_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and there only as the checker / reported
CPU baseline.

PARITY UNPINNED: the reference (XiaRho/MADM) ships no tests, golden vectors or
fixtures for this path, and its arithmetic lives in third-party packages that are
absent here and cannot be installed offline (diffusers==0.25.0, peft==0.10.0,
detectron2@HEAD; reference ``requirements.txt:2,14``, ``README.md:34``).  The oracle
restates their published algorithms (SD-1.4 ``unet/config.json``, ``vae/config.json``,
``scheduler_config.json``) and is pinned by self-made known-answer tests only:
exact SD-1.4 parameter counts (UNet 859,520,964; VAE encoder 34,163,592;
quant_conv 72), the reference's own shape comments
(``modeling/backbone/feature_extractor.py:321-346``), the DDPM alpha-bar closed form
cross-checked against the reference's in-tree ``ldm_linear`` schedule
(``modeling/diffusion/gaussian_diffusion.py:111-121,259-276``), and merged-vs-unmerged
LoRA equality.
"""
