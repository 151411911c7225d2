// Test-infrastructure shim (NOT product code): forward declaration standing in
// for the flatc-generated header.  The reference's api_data structs only hold a
// pointer to gamma_api::EngineStatus; the (de)serialisers that need the full type are
// not part of the oracle build.
#pragma once
namespace gamma_api { struct EngineStatus; }
