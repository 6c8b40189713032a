"""Training path of the backbone (placeholder until the backward engine lands in this round)."""


def extract_with_grad(backbone, img, input_modal, ema_forward, timestep, grad_inputs, **kwargs):
    names = [n for n, _ in grad_inputs]
    raise NotImplementedError(
        f"backbone called under torch.enable_grad() with {len(names)} trainable parameters (e.g. {names[:3]}): "
        "wrap inference in torch.no_grad()")
