"""bnn_priors_b200.overlay against the live reference (build container only): after
install() the reference's runner modules construct the B200 sampler classes."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "bnn_priors")), reason="no reference checkout here")
def test_overlay_rebinds_the_reference_samplers():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(HERE, "golden", "_shims"))
    sys.path.insert(0, REFERENCE)
    try:
        import bnn_priors.mcmc as ref_mcmc
        from bnn_priors import inference, inference_reject
        from bnn_priors_b200 import mcmc as fast, overlay
        original = ref_mcmc.VerletSGLD
        assert original is not fast.VerletSGLD
        overlay.install()
        try:
            # what the runners will construct (inference.py:89-94, inference_reject.py:12-16)
            assert inference.mcmc.SGLD is fast.SGLD
            assert inference.mcmc.HMC is fast.HMC
            assert inference_reject.mcmc.VerletSGLD is fast.VerletSGLD
            from bnn_priors.mcmc.sgld import SGLD as by_submodule
            assert by_submodule is fast.SGLD
        finally:
            overlay.uninstall()
        assert ref_mcmc.VerletSGLD is original and inference.mcmc.SGLD is not fast.SGLD
        # evaluate_model (exp_utils.py:250) as the runners and the experiment scripts reach it
        from bnn_priors import exp_utils
        from bnn_priors_b200.evaluate import evaluate_model as fast_eval
        ref_eval = exp_utils.evaluate_model
        overlay.install(evaluate=True)
        try:
            assert inference.evaluate_model is fast_eval and inference_reject.evaluate_model is fast_eval
            assert exp_utils.evaluate_model is fast_eval
        finally:
            overlay.uninstall()
        assert exp_utils.evaluate_model is ref_eval and inference.evaluate_model is ref_eval
        # fuse_prior / sample_sink: the runner methods and the sample saver are re-bound, and restored
        from bnn_priors_b200.sample_sink import FlatSampleSaver
        make_opt = inference.SGLDRunner.__dict__["_make_optimizer"]
        make_opt_reject = inference_reject.VerletSGLDRunnerReject.__dict__["_make_optimizer"]
        pot = inference.SGLDRunner.__dict__["_model_potential_and_grad"]
        exact = inference_reject.VerletSGLDRunnerReject.__dict__["_exact_model_potential_and_grad"]
        saver = exp_utils.HDF5ModelSaver
        overlay.install(fuse_prior=True, sample_sink=True)
        try:
            assert inference.SGLDRunner.__dict__["_make_optimizer"] is not make_opt
            assert inference_reject.VerletSGLDRunnerReject.__dict__["_make_optimizer"] is not make_opt_reject
            assert inference_reject.HMCRunnerReject.__dict__["_make_optimizer"].__name__ == "_make_optimizer"
            assert inference.SGLDRunner.__dict__["_model_potential_and_grad"] is not pot
            assert inference_reject.VerletSGLDRunnerReject.__dict__["_exact_model_potential_and_grad"] is not exact
            assert exp_utils.HDF5ModelSaver is FlatSampleSaver
            assert issubclass(exp_utils.HDF5Metrics, saver)          # the metrics writer keeps the reference's base class
            overlay.install(fuse_prior=True, sample_sink=True)       # idempotent
        finally:
            overlay.uninstall()
        assert inference.SGLDRunner.__dict__["_make_optimizer"] is make_opt
        assert inference_reject.VerletSGLDRunnerReject.__dict__["_make_optimizer"] is make_opt_reject
        assert inference.SGLDRunner.__dict__["_model_potential_and_grad"] is pot
        assert inference_reject.VerletSGLDRunnerReject.__dict__["_exact_model_potential_and_grad"] is exact
        assert exp_utils.HDF5ModelSaver is saver and ref_mcmc.VerletSGLD is original
        # the submodule bindings too: the reference's own `super(SGLD, self)` (mcmc/sgld.py:39) resolves them
        from bnn_priors.mcmc import sgld as ref_sgld, verlet_sgld as ref_verlet, hmc as ref_hmc
        assert ref_sgld.SGLD is not fast.SGLD and issubclass(original, ref_sgld.SGLD)
        assert ref_verlet.VerletSGLD is original and ref_hmc.HMC is not fast.HMC
    finally:
        sys.path.remove(REFERENCE)
        sys.path.remove(os.path.join(HERE, "golden", "_shims"))
