"""Small models shared by the tests (same architectures as tests/golden/make_golden.py)."""
import torch


class MLP(torch.nn.Module):
    def __init__(self, d_in=20, d_h=16, d_out=5):
        super().__init__()
        self.fc1 = torch.nn.Linear(d_in, d_h)
        self.fc2 = torch.nn.Linear(d_h, d_out)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


class RegNet(torch.nn.Module):
    def __init__(self, n_hid=30):
        super().__init__()
        self.fc1 = torch.nn.Linear(1, n_hid)
        self.fc2 = torch.nn.Linear(n_hid, n_hid)
        self.fc3 = torch.nn.Linear(n_hid, 1)

    def forward(self, x):
        x = torch.relu(self.fc1(x))
        x = torch.relu(self.fc2(x))
        return self.fc3(x)


def load_params(model, golden, prefix, dtype=torch.float64):
    sd = {k: torch.tensor(golden[f"{prefix}_param_{k}"]).to(dtype) for k in model.state_dict().keys()}
    model.to(dtype)
    model.load_state_dict(sd)
    return model
