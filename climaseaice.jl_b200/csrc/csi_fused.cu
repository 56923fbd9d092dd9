// csi_fused.cu -- fused EVP substep (placeholder until the streaming kernel lands).
#include <stdio.h>
#include "csi_internal.h"
namespace csi {
struct FusedPlan { int unused; };
int fused_supported(const DGrid &, const DParams &, const DFields &, char *why, int nwhy) { snprintf(why, nwhy, "not built"); return 0; }
FusedPlan *fused_create(const DGrid &, const DParams &, char *err, int nerr) { snprintf(err, nerr, "not built"); return nullptr; }
void fused_destroy(FusedPlan *p) { delete p; }
int fused_run(FusedPlan *, const LaunchCtx &, const DGrid &, const DParams &, const DFields &, double, int, int, char *err, int nerr) { snprintf(err, nerr, "not built"); return CSI_ERR_UNSUPPORTED; }
}
