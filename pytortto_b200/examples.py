"""Model definitions of the reference's example notebooks, written against the (unchanged) tortto module API.
They are the workloads BASELINE.json names; nothing here is specific to the B200 path - pass any module namespace
with the tortto API (`make_models(tt)`), e.g. this package or the reference itself.

  * PreactResNet / BasicBlock      examples/resnet/preact_resnet18/preact_resnet18.ipynb (cell 10)
    small_preact_resnet110         examples/resnet/small_preact_resnet_110/small_preact_resnet110.ipynb
  * ResNet / Bottleneck (resnet50) examples/resnet/resnet50_finetune/resnet50_finetune.ipynb (cells 152-236)
  * UNet                           examples/unet/UNet.ipynb (cells 175-231)
"""


def make_models(tt):
    nn = tt.nn

    def conv3x3(in_channels, channels, stride=1):
        return nn.Conv2d(in_channels, channels, kernel_size=3, stride=stride, padding=1, bias=False)

    def conv1x1(in_channels, channels, stride=1):
        return nn.Conv2d(in_channels, channels, kernel_size=1, stride=stride, bias=False)

    class BasicBlock(nn.Module):
        expansion = 1

        def __init__(self, in_channels, channels, stride=1, downsample=None):
            super().__init__()
            self.act = nn.Sequential(nn.BatchNorm2d(in_channels), nn.ReLU())
            self.residual = nn.Sequential(conv3x3(in_channels, channels, stride), nn.BatchNorm2d(channels), nn.ReLU(),
                                          conv3x3(channels, channels))
            self.downsample = nn.Sequential() if downsample is None else downsample

        def forward(self, x):
            out = self.act(x)
            shortcut = self.downsample(x)
            out = self.residual(out)
            return out + shortcut

    class PreactResNet(nn.Module):
        def __init__(self, block, layers, channels, num_classes=10):
            super().__init__()
            self.in_channels = channels[0]
            self.conv1 = conv3x3(3, self.in_channels)
            self.layer1 = self._make_layer(block, channels[0], layers[0])
            self.layer2 = self._make_layer(block, channels[1], layers[1], stride=2)
            self.layer3 = self._make_layer(block, channels[2], layers[2], stride=2)
            if len(layers) > 3:
                self.layer4 = self._make_layer(block, channels[3], layers[3], stride=2)
            self.has_layer4 = len(layers) > 3
            self.bn = nn.BatchNorm2d(self.in_channels)
            self.relu = nn.ReLU()
            self.fc = nn.Sequential(nn.Linear(self.in_channels, num_classes), nn.LogSoftmax(dim=-1))

        def _make_layer(self, block, channels, blocks, stride=1):
            downsample = None
            if stride != 1 or self.in_channels != channels * block.expansion:
                downsample = nn.Sequential(conv1x1(self.in_channels, channels * block.expansion, stride))
            layers = [block(self.in_channels, channels, stride, downsample)]
            self.in_channels = channels * block.expansion
            for _ in range(1, blocks):
                layers.append(block(self.in_channels, channels))
            return nn.Sequential(*layers)

        def forward(self, x):
            x = self.conv1(x)
            x = self.layer1(x)
            x = self.layer2(x)
            x = self.layer3(x)
            if self.has_layer4:
                x = self.layer4(x)
            x = self.bn(x)
            x = self.relu(x)
            x = tt.mean(x, (-1, -2), True)
            x = tt.flatten(x, 1)
            return self.fc(x)

    def preact_resnet18(num_classes=10):
        return PreactResNet(BasicBlock, [2, 2, 2, 2], [64, 128, 256, 512], num_classes)

    def small_preact_resnet110(num_classes=10):
        return PreactResNet(BasicBlock, [18, 18, 18], [16, 32, 64], num_classes)

    class Bottleneck(nn.Module):
        expansion = 4

        def __init__(self, in_channels, channels, stride=1, downsample=None):
            super().__init__()
            self.residual = nn.Sequential(
                conv1x1(in_channels, channels), nn.BatchNorm2d(channels), nn.ReLU(),
                conv3x3(channels, channels, stride), nn.BatchNorm2d(channels), nn.ReLU(),
                conv1x1(channels, channels * self.expansion), nn.BatchNorm2d(channels * self.expansion))
            self.downsample = nn.Sequential() if downsample is None else downsample
            self.relu = nn.ReLU()

        def forward(self, x):
            return self.relu(self.residual(x) + self.downsample(x))

    class ResNet(nn.Module):
        def __init__(self, block, layers, channels, num_classes=10):
            super().__init__()
            self.in_channels = 64
            self.stem = nn.Sequential(nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False),
                                      nn.BatchNorm2d(64), nn.ReLU(), nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
            self.layer1 = self._make_layer(block, channels[0], layers[0])
            self.layer2 = self._make_layer(block, channels[1], layers[1], stride=2)
            self.layer3 = self._make_layer(block, channels[2], layers[2], stride=2)
            self.layer4 = self._make_layer(block, channels[3], layers[3], stride=2)
            self.classifier = nn.Sequential(nn.Linear(self.in_channels, num_classes), nn.LogSoftmax(dim=-1))

        def _make_layer(self, block, channels, blocks, stride=1):
            downsample = None
            if stride != 1 or self.in_channels != channels * block.expansion:
                downsample = nn.Sequential(conv1x1(self.in_channels, channels * block.expansion, stride),
                                           nn.BatchNorm2d(channels * block.expansion))
            layers = [block(self.in_channels, channels, stride, downsample)]
            self.in_channels = channels * block.expansion
            for _ in range(1, blocks):
                layers.append(block(self.in_channels, channels))
            return nn.Sequential(*layers)

        def forward(self, x):
            x = self.stem(x)
            x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
            x = tt.mean(x, (-1, -2), True)
            x = tt.flatten(x, 1)
            return self.classifier(x)

    def standard_resnet50(num_classes=10):
        return ResNet(Bottleneck, [3, 4, 6, 3], [64, 128, 256, 512], num_classes)

    class DoubleConv(nn.Module):
        def __init__(self, in_channels, out_channels):
            super().__init__()
            self.conv = nn.Sequential(
                nn.Conv2d(in_channels, out_channels, 3, 1, 1), nn.BatchNorm2d(out_channels), nn.ReLU(),
                nn.Conv2d(out_channels, out_channels, 3, 1, 1), nn.BatchNorm2d(out_channels), nn.ReLU())

        def forward(self, x):
            return self.conv(x)

    class UNet(nn.Module):
        """examples/unet/UNet.ipynb cell 12 (attribute names kept: down, up, pool, bottle_neck, out)"""

        def __init__(self, in_channels, out_channels, features):
            super().__init__()
            self.down = nn.ModuleList()
            self.up = nn.ModuleList()
            self.pool = nn.MaxPool2d(2, 2)
            for f in features:
                self.down.append(DoubleConv(in_channels, f))
                in_channels = f
            self.bottle_neck = DoubleConv(f, 2 * f)
            for f in reversed(features):
                self.up.append(nn.ConvTranspose2d(2 * f, f, kernel_size=2, stride=2))
                self.up.append(DoubleConv(2 * f, f))
            self.out = nn.Conv2d(f, out_channels, kernel_size=1)

        def forward(self, x):
            skip_connections = []
            for m in self.down:
                x = m(x)
                skip_connections.append(x)
                x = self.pool(x)
            x = self.bottle_neck(x)
            skip_connections = skip_connections[::-1]
            for i in range(0, len(self.up), 2):
                skip_connection = skip_connections[i // 2]
                x = self.up[i](x, output_size=skip_connection.shape[-2:])
                x = tt.cat([skip_connections[i // 2], x], dim=1)
                x = self.up[i + 1](x)
            return self.out(x)

    return {"BasicBlock": BasicBlock, "PreactResNet": PreactResNet, "preact_resnet18": preact_resnet18,
            "small_preact_resnet110": small_preact_resnet110, "Bottleneck": Bottleneck, "ResNet": ResNet,
            "standard_resnet50": standard_resnet50, "DoubleConv": DoubleConv, "UNet": UNet}
