"""A second, module-based statement of the RMVPE network (test infrastructure): torch.nn classes laid out like the
published RMVPE model code (ConvBlockRes / ResEncoderBlock / Encoder / Intermediate / ResDecoderBlock / Decoder /
DeepUnet / BiGRU / E2E with n_blocks=4, n_gru=1, kernel_size=(2,2), en_de_layers=5, inter_layers=4, in_channels=1,
en_out_channels=16), so that its state_dict keys ARE the checkpoint's keys.  oracle/nets.py restates the same network
functionally; tests/test_host_cpu.py loads one set of weights into both (strict=True) and compares the salience."""
import torch
import torch.nn as nn


class ConvBlockRes(nn.Module):
    def __init__(self, cin, cout, momentum=0.01):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(cin, cout, (3, 3), (1, 1), (1, 1), bias=False), nn.BatchNorm2d(cout, momentum=momentum), nn.ReLU(),
            nn.Conv2d(cout, cout, (3, 3), (1, 1), (1, 1), bias=False), nn.BatchNorm2d(cout, momentum=momentum), nn.ReLU())
        self.is_shortcut = cin != cout
        if self.is_shortcut:
            self.shortcut = nn.Conv2d(cin, cout, (1, 1))

    def forward(self, x):
        return self.conv(x) + (self.shortcut(x) if self.is_shortcut else x)


class ResEncoderBlock(nn.Module):
    def __init__(self, cin, cout, kernel_size, n_blocks):
        super().__init__()
        self.conv = nn.ModuleList([ConvBlockRes(cin, cout)] + [ConvBlockRes(cout, cout) for _ in range(n_blocks - 1)])
        self.kernel_size = kernel_size
        if kernel_size is not None:
            self.pool = nn.AvgPool2d(kernel_size=kernel_size)

    def forward(self, x):
        for c in self.conv:
            x = c(x)
        return (x, self.pool(x)) if self.kernel_size is not None else x


class Encoder(nn.Module):
    def __init__(self, in_channels, in_size, n_encoders, kernel_size, n_blocks, out_channels=16):
        super().__init__()
        self.bn = nn.BatchNorm2d(in_channels, momentum=0.01)
        self.layers = nn.ModuleList()
        for _ in range(n_encoders):
            self.layers.append(ResEncoderBlock(in_channels, out_channels, kernel_size, n_blocks))
            in_channels, out_channels = out_channels, out_channels * 2
        self.out_channel = out_channels

    def forward(self, x):
        skips = []
        x = self.bn(x)
        for layer in self.layers:
            t, x = layer(x)
            skips.append(t)
        return x, skips


class Intermediate(nn.Module):
    def __init__(self, cin, cout, n_inters, n_blocks):
        super().__init__()
        self.layers = nn.ModuleList([ResEncoderBlock(cin, cout, None, n_blocks)] +
                                    [ResEncoderBlock(cout, cout, None, n_blocks) for _ in range(n_inters - 1)])

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class ResDecoderBlock(nn.Module):
    def __init__(self, cin, cout, stride, n_blocks):
        super().__init__()
        self.conv1 = nn.Sequential(nn.ConvTranspose2d(cin, cout, (3, 3), stride, (1, 1), (1, 1), bias=False),
                                   nn.BatchNorm2d(cout, momentum=0.01), nn.ReLU())
        self.conv2 = nn.ModuleList([ConvBlockRes(cout * 2, cout)] + [ConvBlockRes(cout, cout) for _ in range(n_blocks - 1)])

    def forward(self, x, skip):
        x = torch.cat((self.conv1(x), skip), dim=1)
        for c in self.conv2:
            x = c(x)
        return x


class Decoder(nn.Module):
    def __init__(self, cin, n_decoders, stride, n_blocks):
        super().__init__()
        self.layers = nn.ModuleList()
        for _ in range(n_decoders):
            self.layers.append(ResDecoderBlock(cin, cin // 2, stride, n_blocks))
            cin //= 2

    def forward(self, x, skips):
        for i, layer in enumerate(self.layers):
            x = layer(x, skips[-1 - i])
        return x


class DeepUnet(nn.Module):
    def __init__(self, kernel_size=(2, 2), n_blocks=4, en_de_layers=5, inter_layers=4, in_channels=1, en_out_channels=16):
        super().__init__()
        self.encoder = Encoder(in_channels, 128, en_de_layers, kernel_size, n_blocks, en_out_channels)
        self.intermediate = Intermediate(self.encoder.out_channel // 2, self.encoder.out_channel, inter_layers, n_blocks)
        self.decoder = Decoder(self.encoder.out_channel, en_de_layers, kernel_size, n_blocks)

    def forward(self, x):
        x, skips = self.encoder(x)
        return self.decoder(self.intermediate(x), skips)


class BiGRU(nn.Module):
    def __init__(self, input_features, hidden_features, num_layers):
        super().__init__()
        self.gru = nn.GRU(input_features, hidden_features, num_layers=num_layers, batch_first=True, bidirectional=True)

    def forward(self, x):
        return self.gru(x)[0]


class E2E(nn.Module):
    def __init__(self, n_blocks=4, n_gru=1, kernel_size=(2, 2)):
        super().__init__()
        self.unet = DeepUnet(kernel_size, n_blocks)
        self.cnn = nn.Conv2d(16, 3, (3, 3), padding=(1, 1))
        self.fc = nn.Sequential(BiGRU(3 * 128, 256, n_gru), nn.Linear(512, 360), nn.Dropout(0.25), nn.Sigmoid())

    def forward(self, mel):                       # mel (B, 128, T)
        mel = mel.transpose(-1, -2).unsqueeze(1)  # (B, 1, T, 128)
        x = self.cnn(self.unet(mel)).transpose(1, 2).flatten(-2)
        return self.fc(x)
