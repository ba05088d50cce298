// oracle shim (test infrastructure only): the reference only needs `using namespace boost::assign;` to name a namespace.
#pragma once
namespace boost { namespace assign {} }
