// ADM / EDM U-Net (models/cm/unet.py:523-790) plan builder.
#include "engine.cuh"

namespace dxmi {

void spec_adm(Net& net) { (void)net; }

}  // namespace dxmi
