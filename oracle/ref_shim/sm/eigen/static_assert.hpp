// stand-in (oracle/ref_shim/Eigen/Core): the static assertions of the reference are no-ops here
#include <Eigen/Core>
