// stand-in: serialisation is not exercised by the reference pin
#ifndef KB_SHIM_BOOST_SERIALIZATION
#define KB_SHIM_BOOST_SERIALIZATION
#define BOOST_SERIALIZATION_NVP(x) x
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_CLASS_VERSION(T, N)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)
namespace boost { namespace serialization { class access; template <typename T> T& base_object(T& t) { return t; } } }
#endif
