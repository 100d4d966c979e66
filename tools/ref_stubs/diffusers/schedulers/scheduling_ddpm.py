class DDPMScheduler:
    """Import-only stub; the flow-matching branch never touches the scheduler."""
    def __init__(self, *a, **k):
        pass
