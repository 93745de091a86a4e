"""Learning-rate schedule of neural_network.py:109-125 (config.yaml:57-62; start_lr per SURVEY F8)."""


def lrfn(epoch, start_lr=1e-5, min_lr=1e-5, max_lr=5e-5, rampup_epochs=5, sustain_epochs=0, exp_decay=0.8):
    start_lr, min_lr, max_lr = float(start_lr), float(min_lr), float(max_lr)
    rampup_epochs, sustain_epochs, exp_decay = int(rampup_epochs), int(sustain_epochs), float(exp_decay)
    if epoch < rampup_epochs:
        return (max_lr - start_lr) / rampup_epochs * epoch + start_lr
    elif epoch < rampup_epochs + sustain_epochs:
        return max_lr
    else:
        return (max_lr - min_lr) * exp_decay ** (epoch - rampup_epochs - sustain_epochs) + min_lr
